"""b2d_vec_log_reduce: vec_log across ranks through the C ABI alone (EB:564-598 semantics over all shards) -- the
NCCL all-reduce is issued by the library on a communicator the HOST created (here: with ctypes on the NCCL that
torch ships), no torch.distributed involved.  One rank in-process on any box; two ranks (two processes, one GPU
each) where the box has two GPUs."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _nccl():
    pats = [os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so*") for p in sys.path if p] + ["/usr/lib/x86_64-linux-gnu/libnccl.so*"]
    for pat in pats:
        for f in sorted(glob.glob(pat)):
            L = C.CDLL(f, mode=C.RTLD_GLOBAL)
            L.ncclGetUniqueId.argtypes = [C.POINTER(NcclUniqueId)]
            L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
            L.ncclCommDestroy.argtypes = [C.c_void_p]
            return L
    pytest.skip("no libnccl found")


def _run_steps(vec, steps, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for _ in range(steps):
        vec.step(torch.rand((vec.num_agents, 4), device="cuda", generator=g) * 2 - 1)


def test_single_rank_reduce_equals_plain_vec_log():
    from drone_b200 import capi
    from drone_b200.vec import RaceVec
    L = _nccl()
    uid = NcclUniqueId()
    assert L.ncclGetUniqueId(C.byref(uid)) == 0
    comm = C.c_void_p()
    torch.cuda.set_device(0)
    assert L.ncclCommInitRank(C.byref(comm), 1, uid, 0) == 0
    a, b = RaceVec(4096, seed=3), RaceVec(4096, seed=3)
    a.reset(3), b.reset(3)
    _run_steps(a, 60, 1), _run_steps(b, 60, 1)
    out = (C.c_float * 9)()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(capi.lib().b2d_vec_log_reduce(a.h, out, comm, st))
    want = b.log()
    got = a._log_dict([float(x) for x in out])
    assert got == want and got["n"] > 0
    # accumulators were cleared: a second call reports nothing
    capi.check(capi.lib().b2d_vec_log_reduce(a.h, out, comm, st))
    assert out[8] == 0.0
    # NULL communicator = plain vec_log
    _run_steps(a, 50, 2), _run_steps(b, 50, 2)
    capi.check(capi.lib().b2d_vec_log_reduce(a.h, out, None, st))
    assert a._log_dict([float(x) for x in out]) == b.log()
    L.ncclCommDestroy(comm)
    a.close(), b.close()


def _two_rank_worker(rank, uid_bytes, path):
    import torch as t
    from drone_b200 import capi
    from drone_b200.vec import RaceVec
    t.cuda.set_device(rank)
    L = _nccl()
    uid = NcclUniqueId()
    C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    assert L.ncclCommInitRank(C.byref(comm), 2, uid, rank) == 0
    n = 8192
    vec = RaceVec(n, seed=5, device=f"cuda:{rank}", env_id_base=rank * n)
    vec.reset(5)
    g = t.Generator(device="cpu").manual_seed(11)
    tape = (t.rand((20, 2 * n, 4), generator=g) * 2 - 1)[:, rank * n:(rank + 1) * n].contiguous().to(f"cuda:{rank}")
    for k in range(80):
        vec.step(tape[k % 20])
    out = (C.c_float * 9)()
    capi.check(capi.lib().b2d_vec_log_reduce(vec.h, out, comm, C.c_void_p(t.cuda.current_stream().cuda_stream)))
    np.save(f"{path}.{rank}.npy", np.array([float(x) for x in out], np.float64))
    L.ncclCommDestroy(comm)
    vec.close()


def test_two_rank_reduce_equals_one_big_vec(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from drone_b200.vec import RaceVec
    L = _nccl()
    uid = NcclUniqueId()
    assert L.ncclGetUniqueId(C.byref(uid)) == 0
    path = os.path.join(tmp_path, "log")
    mp.spawn(_two_rank_worker, args=(bytes(uid.internal), path), nprocs=2, join=True)
    got = [np.load(f"{path}.{r}.npy") for r in range(2)]
    assert np.array_equal(got[0], got[1])  # every rank receives the global averages
    n = 8192
    big = RaceVec(2 * n, seed=5)
    big.reset(5)
    g = torch.Generator(device="cpu").manual_seed(11)
    tape = (torch.rand((20, 2 * n, 4), generator=g) * 2 - 1).cuda()
    for k in range(80):
        big.step(tape[k % 20])
    want = big.log()
    assert big._log_dict([float(x) for x in got[0]]) == want  # sharding-invariant: global env ids key the reset stream
    big.close()
