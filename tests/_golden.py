"""Loaders for the committed golden vectors (tests/golden/, minted by make_golden.py from the reference's own code)."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeds():
    with open(os.path.join(GOLDEN, "seeds.json")) as f:
        return json.load(f)


def npz(name):
    return np.load(os.path.join(GOLDEN, name))


def cases(z):
    return json.loads(str(z["cases"]))


def meta(z):
    return json.loads(str(z["meta"]))


def mds42():
    out = []
    for name in ("mds42_recoded", "mds42_full"):
        with gzip.open(os.path.join(GOLDEN, name + ".fa.gz"), "rb") as f:
            lines = f.read().split(b"\n")
        out.append(b"".join(l for l in lines if l and not l.startswith(b">")))
    return out


def canon_ties(pos, mer):
    return np.asarray(pos)[np.lexsort((pos, mer))]
