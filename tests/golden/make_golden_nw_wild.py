"""Mints tests/golden/nw_wild.npz: NWSmall + BitTraceBack paths of the REFERENCE's own code (oracle/_ref/libmauve_ref.so: ProfileProfile ->
GlobalAlign -> NWSmall) for sequence pairs that contain DNA wildcards (N, X, R, Y, ... either case), the inputs mcu_nw_batch_wild covers.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_nw_wild.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402


def cases():
    rng = np.random.default_rng(20261022)
    out = []
    for wild in (b"N", b"NnXx", b"MRWSYKVHDBXNmrwsykvhdbxn", b"acgtN"):
        for _ in range(45):
            a = bytes(rng.choice(list(b"ACGT") * 6 + list(wild), int(rng.integers(1, 260))).astype(np.uint8))
            if rng.random() < 0.7:   # a diverged copy with small indels
                s = bytearray(a)
                for i in rng.integers(0, len(s), max(1, len(s) // 10)):
                    s[i] = int(rng.choice(list(b"ACGT") + list(wild)))
                k = int(rng.integers(0, len(s)))
                del s[k:k + int(rng.integers(0, min(8, len(s))))]
                k = int(rng.integers(0, len(s) + 1))
                s[k:k] = bytes(rng.choice(list(b"ACGT") + list(wild), int(rng.integers(0, 8))).astype(np.uint8))
                b = bytes(s) or b"N"
            else:
                b = bytes(rng.choice(list(b"ACGT") * 6 + list(wild), int(rng.integers(1, 260))).astype(np.uint8))
            out.append((a, b))
    out += [(b"N", b"N"), (b"X", b"A"), (b"A", b"NNNNNNNN"), (b"NNNNNNNN", b"G"), (b"ACGTNACGT" * 30, b"ACGTACGT" * 33), (b"n" * 100, b"x" * 90)]
    return out


def main():
    ref = _oracle.ref_checker()
    data = {}
    cs = cases()
    for i, (a, b) in enumerate(cs):
        p, _ = ref.nw_align(a, b)
        data["a%d" % i] = np.frombuffer(a, dtype=np.uint8)
        data["b%d" % i] = np.frombuffer(b, dtype=np.uint8)
        data["p%d" % i] = np.frombuffer(p, dtype=np.uint8)
    data["n"] = np.int64(len(cs))
    np.savez_compressed(os.path.join(HERE, "nw_wild.npz"), **data)
    print("nw_wild.npz: %d pairs" % len(cs))


if __name__ == "__main__":
    main()
