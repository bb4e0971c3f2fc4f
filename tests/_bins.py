"""TEST INFRASTRUCTURE ONLY: the binaries oracle/Makefile.ref builds under oracle/_ref and small helpers to drive them."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
NEXT_BIN = os.path.join(REF_DIR, "dropin_check_next")   # drop-in check of the rows next to the hot path (all of libMems)


def run_kv(binary, args, env=None, timeout=900):
    """run a check binary; returns (exit code, {key: value} of its "key value" output lines, stdout + stderr)"""
    r = subprocess.run([binary] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout, env=env)
    kv = dict(l.split(" ", 1) for l in r.stdout.splitlines() if " " in l)
    return r.returncode, kv, r.stdout + r.stderr


def run_next(args, env=None, timeout=900):
    return run_kv(NEXT_BIN, args, env, timeout)


def write_fasta(path, name, seq, width=80):
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), width):
            f.write(seq[i:i + width] + b"\n")
