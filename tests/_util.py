"""Helpers shared by the parity tests (test infrastructure)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def row_hash(*cols):
    """FNV-1a over the 32-bit words of each row: one uint32 per row.  `cols` are arrays
    with the same leading shape; float32 is hashed by bit pattern, so equal hashes mean
    bit-identical rows (up to 2^-32 collisions)."""
    words = []
    for c in cols:
        c = np.asarray(c)
        if c.dtype == np.float32:
            c = c.view(np.uint32)
        else:
            c = c.astype(np.uint32)
        words.append(c.reshape(c.shape[0], -1))
    w = np.concatenate(words, axis=1).astype(np.uint64)
    h = np.full(w.shape[0], 2166136261, np.uint64)
    for k in range(w.shape[1]):
        h = ((h ^ w[:, k]) * np.uint64(16777619)) & np.uint64(0xFFFFFFFF)
    return h.astype(np.uint32)


def action_tape(n, steps=16, seed=1234, scale=1.3):
    rng = np.random.default_rng(seed)
    return rng.uniform(-scale, scale, size=(steps, n, 4)).astype(np.float32)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))
