"""The advantage oracle (oracle/drone_oracle.c orc_puff_advantage = pufferlib.cpp:28-41,63-72).

PINNED against the reference itself: tests/golden/advantage_*.npz hold outputs of the UNMODIFIED
pufferlib/extensions/pufferlib.cpp (compiled into oracle/_ref/libref_advantage.so by oracle/Makefile, vectors
written by tests/golden/make_golden_advantage.py); the restatement must reproduce them bit for bit, and where
oracle/_ref is present the reference is also run live on fresh inputs.  (The reference's own test for this op,
tests/test_c_advantage.cu, is stale and does not build)  An independent float64 statement
of the recurrence and a literal float32 transcription stay as cross-checks."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "advantage_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_restatement_reproduces_the_reference_golden_vectors(oracle, path):
    g = np.load(path)
    gamma, lam, rho, c = (float(x) for x in g["hyper"])
    adv, prio = oracle.puff_advantage(g["values"], g["rewards"], g["dones"], g["importance"], gamma, lam, rho, c)
    assert np.array_equal(adv.view(np.uint32), g["advantages"].view(np.uint32))
    tm, _ = oracle.puff_advantage(g["values"].T.copy(), g["rewards"].T.copy(), g["dones"].T.copy(), g["importance"].T.copy(),
                                  gamma, lam, rho, c, time_major=True)
    assert np.array_equal(tm.T.copy().view(np.uint32), g["advantages"].view(np.uint32))


def test_golden_files_exist():
    assert len(GOLDEN) >= 4


def test_restatement_equals_the_live_reference(oracle):
    """Fresh inputs through the compiled reference source (build container / any box that carries oracle/_ref)."""
    if not oracle.have_ref_advantage():
        pytest.skip("oracle/_ref/libref_advantage.so not present")
    for seed, (rows, horizon) in enumerate([(1, 2), (257, 64), (33, 1), (1000, 128)]):
        v, r, d, imp = _inputs(rows, horizon, seed=100 + seed)
        for rho, c in ((1.0, 1.0), (0.8, 1.2)):
            want = oracle.ref_puff_advantage(v, r, d, imp, 0.99, 0.95, rho, c)
            adv, _ = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, rho, c)
            assert np.array_equal(adv.view(np.uint32), want.view(np.uint32)), (rows, horizon, rho, c)


def _inputs(rows, horizon, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.normal(0, 1, (rows, horizon)).astype(np.float32)
    r = np.clip(rng.normal(0, 0.5, (rows, horizon)), -1, 1).astype(np.float32)
    d = (rng.random((rows, horizon)) < 0.05).astype(np.float32)
    imp = np.exp(rng.normal(0, 0.3, (rows, horizon))).astype(np.float32)
    return v, r, d, imp


def _literal_f32(v, r, d, imp, gamma, lam, rho_clip, c_clip):
    f = np.float32
    adv = np.zeros_like(v)
    for row in range(v.shape[0]):
        last = f(0)
        for t in range(v.shape[1] - 2, -1, -1):
            nnt = f(1.0 - float(d[row, t + 1]))
            rho = min(imp[row, t], f(rho_clip))
            c = min(imp[row, t], f(c_clip))
            delta = f(rho * f(f(f(r[row, t + 1] + f(f(f(gamma) * v[row, t + 1]) * nnt)) - v[row, t])))
            last = f(delta + f(f(f(f(f(gamma) * f(lam)) * c) * last) * nnt))
            adv[row, t] = last
    return adv


def test_oracle_matches_literal_float32_transcription(oracle):
    v, r, d, imp = _inputs(24, 17)
    adv, prio = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, 1.0, 1.0)
    want = _literal_f32(v, r, d, imp, 0.99, 0.95, 1.0, 1.0)
    assert np.array_equal(adv.view(np.uint32), want.view(np.uint32))
    assert np.all(adv[:, -1] == 0)  # the last column is never written (horizon-2 .. 0)
    assert np.allclose(prio, np.abs(adv).sum(1), rtol=1e-5)


@pytest.mark.parametrize("rho_clip,c_clip", [(1.0, 1.0), (0.7, 1.3), (100.0, 100.0)])
def test_oracle_matches_float64_recurrence(oracle, rho_clip, c_clip):
    v, r, d, imp = _inputs(64, 64, seed=2)
    adv, _ = oracle.puff_advantage(v, r, d, imp, 0.995, 0.9, rho_clip, c_clip)
    V, R, D, I = (x.astype(np.float64) for x in (v, r, d, imp))
    want = np.zeros_like(V)
    last = np.zeros(V.shape[0])
    for t in range(V.shape[1] - 2, -1, -1):
        nnt = 1.0 - D[:, t + 1]
        delta = np.minimum(I[:, t], rho_clip) * (R[:, t + 1] + 0.995 * V[:, t + 1] * nnt - V[:, t])
        last = delta + 0.995 * 0.9 * np.minimum(I[:, t], c_clip) * last * nnt
        want[:, t] = last
    assert np.allclose(adv, want, rtol=2e-5, atol=2e-5)


def test_oracle_time_major_equals_row_major(oracle):
    v, r, d, imp = _inputs(40, 33, seed=3)
    a, pa = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, 1.0, 1.0)
    b, pb = oracle.puff_advantage(v.T.copy(), r.T.copy(), d.T.copy(), imp.T.copy(), 0.99, 0.95, 1.0, 1.0, time_major=True)
    assert np.array_equal(a.view(np.uint32), b.T.copy().view(np.uint32)) and np.array_equal(pa, pb)


def test_oracle_degenerate_horizons(oracle):
    for horizon in (1, 2):
        v, r, d, imp = _inputs(5, horizon, seed=4)
        adv, prio = oracle.puff_advantage(v, r, d, imp, 0.99, 0.95, 1.0, 1.0)
        assert adv.shape == (5, horizon) and np.all(adv[:, -1] == 0)
