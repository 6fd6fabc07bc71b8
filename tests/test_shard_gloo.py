"""World-size-2 test of the multi-GPU host logic on CPU (gloo): env sharding by global env id
and the vec_log reduction.  Each rank drives the CPU oracle for its shard (test
infrastructure standing in for the device), all-reduces its integer episode sums with
drone_b200.shard.reduce_log_sums, averages them through the C ABI (b2d_log_average) and the
result must equal the unsharded run's vec_log.  Observations must be identical to the
corresponding rows of the unsharded run (reset streams keyed by global env id).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, T, SEED, MAX_RINGS, MAX_MOVES = 96, 160, 21, 10, 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(po, lo, n, tape):
    """Returns (obs hash of the last step, int64 sums[16] in b2d_vec_log_begin order)."""
    env = po.OrcRace(n, max_rings=MAX_RINGS, max_moves=MAX_MOVES, seed=SEED, env_id_base=lo)
    env.reset(SEED, mode=po.RESET_PHILOX)
    sums = np.zeros(16, np.int64)
    rings = np.zeros(n, np.int64)
    ticks = np.zeros(n, np.int64)
    rets = np.zeros(n, np.int64)
    for t in range(T):
        env.step(tape[t % 16][lo:lo + n], mode=po.RESET_PHILOX)
        ev = env.events
        ticks += 1
        rings += (ev & po.EV_RING_PASS) != 0
        rets += env.rewards.astype(np.int64)
        done = env.terminals == 1
        last = 0
        if done.any():
            sums[0] += done.sum()                      # ACC_N
            sums[1] += rets[done].sum()                # ACC_RETURN
            sums[2] += ticks[done].sum()               # ACC_LENGTH
            sums[3] += rings[done].sum()               # ACC_RINGS
            sums[4] += ((ev & po.EV_OOB) != 0).sum()   # ACC_OOB
            sums[5] += ((ev & po.EV_COLLISION) != 0).sum()
            sums[6] += ((ev & po.EV_TIMEOUT) != 0).sum()
            last = rings[done].sum()
            rings[done] = 0
            ticks[done] = 0
            rets[done] = 0
        sums[8] = last                                  # score of the last step only (drone_race.h:160)
    obs = env.observations.copy()
    log = env.log()
    env.close()
    return obs, sums, log


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from drone_b200 import shard
    from oracle import pyoracle as po
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, n = shard.shard_range(N_TOTAL, rank, world)
    tape = np.random.default_rng(5).uniform(-1, 1, (16, N_TOTAL, 4)).astype(np.float32)
    obs, sums, _ = _run_shard(po, lo, n, tape)
    tsum = torch.from_numpy(sums.copy())
    shard.reduce_log_sums(tsum)
    avg = shard.average_log(tsum.tolist(), shard.KIND_RACE, MAX_RINGS)
    q.put((rank, lo, n, obs, avg))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from drone_b200 import shard
    for total, world in [(96, 2), (1 << 20, 8), (10, 3), (7, 8)]:
        parts = [shard.shard_range(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(n for _, n in parts) == total
        for (lo, n), (lo2, _) in zip(parts, parts[1:]):
            assert lo + n == lo2
    assert shard.env_id_base(1 << 20, 3) == 3 << 20
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_average_log_matches_env_binding_semantics():
    from drone_b200 import shard
    sums = [0] * 16
    assert shard.average_log(sums) == [0.0] * 9  # reference returns {} when n == 0 (EB:582-585)
    sums[0], sums[1], sums[2], sums[3], sums[4], sums[5], sums[6], sums[8] = 8, -6, 400, 4, 5, 2, 1, 3
    out = shard.average_log(sums, shard.KIND_RACE, 10)
    assert out[8] == 8.0 and out[0] == -0.75 and out[1] == 50.0
    assert out[4] == 0.625 and out[3] == 0.25 and out[5] == 0.125
    assert out[7] == pytest.approx(4 / 10 / 8) and out[6] == 0.375


def test_two_rank_gloo_log_reduction_and_shard_invariance():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from drone_b200 import shard
    from oracle import pyoracle as po
    tape = np.random.default_rng(5).uniform(-1, 1, (16, N_TOTAL, 4)).astype(np.float32)
    obs_all, sums_all, log_all = _run_shard(po, 0, N_TOTAL, tape)
    assert sums_all[0] > 50
    want = shard.average_log(sums_all.tolist(), shard.KIND_RACE, MAX_RINGS)
    # the integer sums reproduce the reference-style float log of the unsharded oracle
    n_ep = float(log_all[8])
    assert want[8] == n_ep
    assert want[1] == pytest.approx(log_all[1] / n_ep, rel=1e-6)
    assert want[0] == pytest.approx(log_all[0] / n_ep, rel=1e-6)
    assert want[4] == pytest.approx(log_all[4] / n_ep, rel=1e-6)
    assert want[7] == pytest.approx(log_all[7] / n_ep, rel=1e-5, abs=1e-7)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, n, obs, avg in res:
        assert np.array_equal(obs.view(np.uint32), obs_all[lo:lo + n].view(np.uint32)), "shard rows differ from the unsharded run"
        assert avg == want, "all-reduced vec_log differs from the unsharded vec_log"
