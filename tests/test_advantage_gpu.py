"""b2d_puff_advantage on the GPU against the reference's golden vectors (tests/golden/advantage_*.npz, outputs of
the unmodified pufferlib/extensions/pufferlib.cpp) and the oracle restatement pinned to them."""
import os

import numpy as np
import pytest

from test_advantage_cpu import GOLDEN, _inputs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,horizon", [(1, 2), (257, 64), (5000, 128), (33, 1)])
@pytest.mark.parametrize("time_major", [False, True])
def test_strict_bit_exact_both_layouts(oracle, rows, horizon, time_major):
    from drone_b200.advantage import compute_puff_advantage
    v, r, d, imp = _inputs(rows, horizon, seed=rows)
    if time_major:
        v, r, d, imp = (np.ascontiguousarray(x.T) for x in (v, r, d, imp))
    want, prio = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, 0.8, 1.2, time_major=time_major)
    tv, tr, td, ti = (torch.from_numpy(x).cuda() for x in (v, r, d, imp))
    adv = torch.zeros_like(tv)
    pr = torch.zeros(rows, device="cuda")
    out = compute_puff_advantage(tv, tr, td, ti, adv, 0.99, 0.95, 0.8, 1.2, time_major=time_major, priority=pr, math="strict")
    assert out is adv
    assert np.array_equal(adv.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(pr.cpu().numpy().view(np.uint32), prio.view(np.uint32))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
@pytest.mark.parametrize("time_major", [False, True])
def test_strict_kernel_reproduces_the_reference_golden_vectors(path, time_major):
    from drone_b200.advantage import compute_puff_advantage
    g = np.load(path)
    gamma, lam, rho, c = (float(x) for x in g["hyper"])
    arrs = [g[k] for k in ("values", "rewards", "dones", "importance")]
    if time_major:
        arrs = [np.ascontiguousarray(x.T) for x in arrs]
    tv, tr, td, ti = (torch.from_numpy(x).cuda() for x in arrs)
    adv = torch.zeros_like(tv)
    compute_puff_advantage(tv, tr, td, ti, adv, gamma, lam, rho, c, time_major=time_major, math="strict")
    got = adv.cpu().numpy()
    if time_major:
        got = np.ascontiguousarray(got.T)
    assert np.array_equal(got.view(np.uint32), g["advantages"].view(np.uint32))


def test_fast_math_within_tolerance_and_errors(oracle):
    from drone_b200.advantage import compute_puff_advantage
    v, r, d, imp = _inputs(4096, 128, seed=1)
    want, _ = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, 1.0, 1.0)
    tv, tr, td, ti = (torch.from_numpy(x).cuda() for x in (v, r, d, imp))
    adv = torch.zeros_like(tv)
    compute_puff_advantage(tv, tr, td, ti, adv, 0.99, 0.95, 1.0, 1.0)
    assert np.allclose(adv.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        compute_puff_advantage(tv.double(), tr, td, ti, adv, 0.99, 0.95, 1.0, 1.0)
    with pytest.raises(ValueError):
        compute_puff_advantage(tv[:, :5], tr, td, ti, adv, 0.99, 0.95, 1.0, 1.0)
    with pytest.raises(ValueError):
        compute_puff_advantage(tv.cpu(), tr, td, ti, adv, 0.99, 0.95, 1.0, 1.0)


def test_advantage_of_a_device_rollout():
    """End of the rollout row: experience collected on the device goes straight into the advantage
    kernel in its time-major layout (no transpose, no host copy)."""
    from drone_b200.advantage import compute_puff_advantage
    from drone_b200.rollout import DeviceRollout, DronePolicy
    from drone_b200.vec import RaceVec
    torch.manual_seed(0)
    vec = RaceVec(4096, seed=1)
    vec.reset(1)
    ro = DeviceRollout(vec, DronePolicy().cuda(), horizon=32).collect()
    adv = torch.zeros_like(ro.values)
    prio = torch.zeros(4096, device="cuda")
    compute_puff_advantage(ro.values, ro.rewards, ro.terminals, torch.ones_like(ro.values), adv, 0.99, 0.95, 1.0, 1.0,
                           time_major=True, priority=prio)
    torch.cuda.synchronize()
    assert torch.isfinite(adv).all() and float(adv.abs().sum()) > 0
    assert torch.allclose(prio, adv.abs().sum(0), rtol=1e-4, atol=1e-5)
    vec.close()
