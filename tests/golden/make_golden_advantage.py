"""Generates tests/golden/advantage_*.npz from the UNMODIFIED reference implementation of
compute_puff_advantage (pufferlib/extensions/pufferlib.cpp:28-41,63-72), compiled into
oracle/_ref/libref_advantage.so by oracle/Makefile.

    python tests/golden/make_golden_advantage.py          # needs /root/reference (build container only)

Each file holds the inputs (values, rewards, dones, importance: float32 [rows, horizon]), the hyper-parameters
and the reference's advantages.  The reference's own test for this op is stale and does not build (SURVEY.md
section 4), so these files are the pin for SURVEY 8f-2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as po  # noqa: E402


def inputs(rows, horizon, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(0, 1, (rows, horizon)).astype(np.float32)
    r = np.clip(rng.normal(0, 0.5, (rows, horizon)), -1, 1).astype(np.float32)
    d = (rng.random((rows, horizon)) < 0.05).astype(np.float32)
    imp = np.exp(rng.normal(0, 0.3, (rows, horizon))).astype(np.float32)
    return v, r, d, imp


CASES = [  # name, rows, horizon, seed, gamma, lambda, rho_clip, c_clip
    ("advantage_r64_h128_default", 64, 128, 11, 0.99, 0.95, 1.0, 1.0),
    ("advantage_r37_h64_clips", 37, 64, 12, 0.995, 0.9, 0.7, 1.3),
    ("advantage_r5_h2_edge", 5, 2, 13, 0.99, 0.95, 1.0, 1.0),
    ("advantage_r3_h1_edge", 3, 1, 14, 0.99, 0.95, 1.0, 1.0),
]

if __name__ == "__main__":
    assert po.have_ref_advantage(), "build oracle/_ref first (make -C oracle)"
    for name, rows, horizon, seed, gamma, lam, rho, c in CASES:
        v, r, d, imp = inputs(rows, horizon, seed)
        adv = po.ref_puff_advantage(v, r, d, imp, gamma, lam, rho, c)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), values=v, rewards=r, dones=d, importance=imp, advantages=adv,
                            hyper=np.array([gamma, lam, rho, c], np.float64))
        print(name, adv.shape, float(np.abs(adv).sum()))
