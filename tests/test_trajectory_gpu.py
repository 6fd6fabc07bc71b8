"""Env-state checkpoint / render bridge on a live device env: b2d_export_ref hands one device env
out as the reference's own structs; the reference's compute_observations on them equals the row the kernel wrote.
Trajectory dumps replay: restoring a recorded state blob and re-applying the recorded actions reproduces the run."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_device_env_through_reference_structs(oracle):
    from drone_b200 import trajectory
    from drone_b200.vec import RaceVec, SwarmVec
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    fp = C.POINTER(C.c_float)
    L = C.CDLL(oracle.REF_RACE_SO)
    vec = RaceVec(300, seed=4, math="strict")
    vec.reset(4)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(25):
        vec.step(torch.rand((300, 4), device="cuda", generator=g) * 2 - 1)
    obs = vec.observations.cpu().numpy()
    for i in (0, 17, 299):
        drones, rings, tick, ring_idx = trajectory.export_ref(vec, i)
        row = np.zeros(29, np.float32)
        L.refrace_observe_structs(C.byref(drones[0]), rings, vec.max_rings, ring_idx, row.ctypes.data_as(fp))
        assert np.array_equal(_bits(row), _bits(obs[i])), i
    vec.close()
    S = C.CDLL(oracle.REF_SWARM_SO)
    sw = SwarmVec(12, 8, 5, seed=2, math="strict")
    sw.reset(2)
    for t in range(10):
        sw.step(torch.rand((96, 4), device="cuda", generator=g) * 2 - 1)
    sobs = sw.observations.cpu().numpy()
    for e in (0, 11):
        drones, rings, tick, task = trajectory.export_ref(sw, e)
        rows = np.zeros((8, 41), np.float32)
        S.refswarm_observe_structs(drones, 8, rings, 5, task, rows.ctypes.data_as(fp))
        assert np.array_equal(_bits(rows), _bits(sobs[e * 8:(e + 1) * 8])), e
    sw.close()


def test_recorded_trajectory_replays(tmp_path):
    from drone_b200 import trajectory
    from drone_b200.vec import RaceVec
    n, T = 512, 40
    g = torch.Generator(device="cuda").manual_seed(1)
    tape = torch.rand((T, n, 4), device="cuda", generator=g) * 2 - 1
    vec = RaceVec(n, seed=7, math="strict")
    vec.reset(7)
    ids = [3, 100, 511]
    path = os.path.join(tmp_path, "traj.npz")
    rec = trajectory.record(vec, T, ids, lambda t: tape[t], path=path)
    vec.close()
    d = np.load(path)
    assert d["states"].shape == (T + 1, 3, 33 + 60) and d["actions"].shape == (T, 3, 4)
    assert np.array_equal(d["states"], rec["states"])
    # replay the first 10 steps of env 100 from its recorded state on a fresh vec: same states while no reset intervenes
    other = RaceVec(n, seed=7, math="strict")
    other.reset(7)
    other.put_state(d["states"][0, 1:2], env_ids=[100])
    for t in range(10):
        if d["terminals"][t, 1]:
            break
        a = torch.zeros((n, 4), device="cuda")
        a[100] = torch.from_numpy(d["actions"][t, 1]).cuda()
        other.step(a)
        got = other.get_state([100])
        if not d["terminals"][t, 1]:
            assert np.array_equal(_bits(got[0, :33]), _bits(d["states"][t + 1, 1, :33])), t
    other.close()
