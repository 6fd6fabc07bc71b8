"""The advantage oracle (oracle/drone_oracle.c orc_puff_advantage = pufferlib.cpp:28-41,63-72) against
an independent float64 statement of the same recurrence and a literal float32 Python transcription.
The reference's own test for this op (tests/test_c_advantage.cu) is stale and does not build
(SURVEY.md section 4): parity for this row is pinned by these two checks only."""
import numpy as np
import pytest


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
