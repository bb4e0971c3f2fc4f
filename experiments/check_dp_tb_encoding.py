#!/usr/bin/env python
"""Exactness check (CPU, numpy) for r02_dp_predicated_traceback_bits.patch: the re-encoded traceback nibble decodes to the same
predecessor state and the same D/I origin bits as the nibble dp.cu writes today, ties included.  It checks the LOGIC of the two
encodings on random and exhaustive small values; the inline PTX of the patch itself only runs on a GPU."""
import itertools

import numpy as np


def old_nibble(M, D, I, upD, upM, Ml, Il):
    best = np.maximum(M, np.maximum(D, I))
    x = np.where(M == best, 0, np.where(D == best, 1, 2))
    return x | np.where(upD > upM, 0, 4) | np.where(Ml >= Il, 8, 0)


def new_nibble(M, D, I, upD, upM, Ml, Il):
    DI = np.maximum(D, I)
    return (DI > M).astype(np.int64) | ((I > D).astype(np.int64) << 1) | ((upM >= upD).astype(np.int64) << 2) | ((Ml >= Il).astype(np.int64) << 3)


def decode_new(nb):
    x = (nb & 1) + (nb & (nb >> 1) & 1)
    return x, nb & 4, nb & 8


def check(upD, upM, Ml, Il, M):
    D = np.maximum(upD, upM)
    I = np.maximum(Ml, Il)
    o = old_nibble(M, D, I, upD, upM, Ml, Il)
    x, b2, b3 = decode_new(new_nibble(M, D, I, upD, upM, Ml, Il))
    assert np.array_equal(x, o & 3) and np.array_equal(b2, o & 4) and np.array_equal(b3, o & 8)


vals = np.array(list(itertools.product(range(-2, 3), repeat=5)), dtype=np.int64)   # every tie pattern
check(*vals.T)
rng = np.random.default_rng(1)
for scale in (3, 500, 1 << 22):
    check(*rng.integers(-scale, scale, size=(5, 1 << 20)))
print("traceback nibble encodings agree (exhaustive [-2,2]^5 and 3 x 2^20 random tuples)")
