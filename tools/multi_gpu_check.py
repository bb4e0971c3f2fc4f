#!/usr/bin/env python
"""Parity + timing of the sharded seed+match+extend step on REAL ranks (one process per GPU, the library's own NCCL layer).

    python tools/multi_gpu_check.py --gpus 2 [--mbp 5] [--steps 5]        # launcher: spawns the ranks itself
    (each rank re-enters this file with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT set, exactly what torchrun provides)

Every rank: whole pair resident (mcu_session_upload), `steps` sharded runs (mcu_session_run_sharded), then the host-buffer
collective (mcu_find_mums_sharded).  Rank 0 also runs the pair unsharded on its own GPU and compares sha1 of the rows; at
--mbp 100 additionally against the reference's own list (tests/golden/config3_rows.json).  Exit code 0 only if everything agrees.
Prints one JSON line (rank 0).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sha(rows):
    import numpy as np
    return hashlib.sha1(np.ascontiguousarray(rows, dtype=np.int64).tobytes()).hexdigest()


def worker(args):
    import numpy as np
    import mauve_py_b200 as mp
    from mauve_py_b200 import dist as mdist, synth
    from mauve_py_b200._capi import check
    comm = mdist.init_from_env()
    rank, world = comm.rank, comm.world
    lib = mp.lib()
    if args.mbp == 5:
        a, b = synth.config2_pair()
    else:
        a, b = synth.config3_pair(n=int(args.mbp * 1e6))
    weight = args.weight or mp.getDefaultSeedWeight((a.size + b.size) // 2)
    seed = mp.getSeed(weight, mp.CODING_SEED)
    ab, bb = a.tobytes(), b.tobytes()
    sess = mp.AnchorSession()
    sess.upload(ab, bb)
    out = {"gpus": world, "mbp": args.mbp, "seed": hex(seed)}
    ms, dev = [], []
    n = 0
    for i in range(args.steps + 2):
        comm.barrier()
        t0 = time.perf_counter()
        n = sess.run_sharded(seed)
        dt = 1e3 * (time.perf_counter() - t0)
        dt = comm.allreduce([dt], mdist.MAX)[0]
        if i >= 2:
            ms.append(dt)
            dev.append(float(sess.stage_ms[6]))
    stage = [round(float(x), 3) for x in sess.stage_ms]
    all_stage = comm.gather_bytes(np.asarray(stage, dtype=np.float64).tobytes())
    rows = sess.download().copy() if rank == 0 else None
    # host-buffer collective
    cap = n + 16
    buf = np.zeros((cap, 3), dtype=np.int64)
    n_out = C.c_uint64(0)
    e2e = []
    for i in range(3):
        comm.barrier()
        t0 = time.perf_counter()
        check(lib.mcu_find_mums_sharded(ab, len(ab), bb, len(bb), seed, 0, buf.ctypes.data, cap, C.byref(n_out), None))
        e2e.append(comm.allreduce([1e3 * (time.perf_counter() - t0)], mdist.MAX)[0])
    # sorted mer list of genome 0 sharded by mer range (mcu_sml_build_sharded): positions gathered on rank 0
    sml_ms = []
    for i in range(3):
        comm.barrier()
        t0 = time.perf_counter()
        spos, dev_ms = mp.libmems.sml_build_sharded(ab, seed)
        sml_ms.append((comm.allreduce([1e3 * (time.perf_counter() - t0)], mdist.MAX)[0], comm.allreduce([dev_ms], mdist.MAX)[0]))
    ok = True
    if rank == 0:
        sml = mp.DNAMemorySML()
        t0 = time.perf_counter()
        sml.Create(ab, seed)
        t1 = 1e3 * (time.perf_counter() - t0)
        pos_u, mer_u = sml.positions(), sml.mers()
        mer_by_pos = np.empty_like(mer_u)
        mer_by_pos[pos_u] = mer_u
        out["sml_sharded"] = {"wall_ms": min(x[0] for x in sml_ms), "device_ms": min(x[1] for x in sml_ms), "single_gpu_wall_ms_with_mers_back": t1,
                              "same_mer_sequence": bool(spos.size == pos_u.size and np.array_equal(mer_by_pos[spos], mer_u)),
                              "a_permutation": bool(np.array_equal(np.sort(spos), np.arange(pos_u.size, dtype=np.uint32)))}
        ok = out["sml_sharded"]["same_mer_sequence"] and out["sml_sharded"]["a_permutation"]
        out.update(rows=int(rows.shape[0]), sharded_ms=sorted(ms), device_ms_rank0=sorted(dev), e2e_ms=sorted(e2e), sha1=sha(rows),
                   stage_ms_per_rank=[np.frombuffer(x, dtype=np.float64).tolist() for x in all_stage])
        out["e2e_same"] = int(n_out.value) == rows.shape[0] and sha(buf[:int(n_out.value)]) == out["sha1"]
        single = mp.AnchorSession()
        single.upload(ab, bb)
        t0 = time.perf_counter()
        n1 = single.run(seed)
        out["single_ms_first"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        single.run(seed)
        out["single_ms"] = 1e3 * (time.perf_counter() - t0)
        srows = single.download().copy()
        out["single_rows"] = int(n1)
        out["same_as_single_gpu"] = bool(srows.shape == rows.shape and sha(srows) == out["sha1"])
        ok = ok and out["same_as_single_gpu"] and out["e2e_same"]
        gp = os.path.join(ROOT, "tests", "golden", "config3_rows.json")
        if args.mbp == 100 and os.path.exists(gp) and not args.weight:
            g = json.load(open(gp))
            out["same_as_reference_golden"] = out["sha1"] == g["reference"]["sha1"]
            ok = ok and out["same_as_reference_golden"]
        print(json.dumps(out), flush=True)
    flag = comm.allreduce([0.0 if ok else 1.0], mdist.MAX)[0]
    comm.barrier()
    comm.close()
    sys.exit(0 if flag == 0.0 else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--mbp", type=float, default=5)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--weight", type=int, default=0)
    ap.add_argument("--port", type=int, default=29591)
    args = ap.parse_args()
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        worker(args)
        return
    procs = []
    for r in range(args.gpus):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(args.gpus), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(args.port),
                   MCU_RENDEZVOUS_TAG="mgc%d" % os.getpid())
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))
    rc = 0
    for p in procs:
        try:
            rc |= p.wait(timeout=900)
        except subprocess.TimeoutExpired:
            p.kill()
            rc |= 124
    sys.exit(rc)


if __name__ == "__main__":
    main()
