#!/usr/bin/env python
"""bench.py -- headline benchmark of the anchoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbp 100]

One "step" = one pass of seed + match + extend (pack -> two partition passes over 8-byte seed records -> in-bucket
grouping -> candidates -> extension -> reference list order) over the synthetic 100 Mbp pair (BASELINE config 3, SURVEY.md 8d
"C3"; the north_star target workload, it fits one GPU).  `value` is Mbp/s with both genomes already resident in HBM; `e2e` is
the same metric through the public C-ABI call (mcu_find_mums_into / mcu_find_mums_sharded) with pinned HOST buffers, H2D + D2H
inside the timed region.  N > 1 shards the SAME pair ("strong" scaling): the whole sharded step is ONE library call
(mcu_session_run_sharded, csrc/comm.cu) that calls NCCL itself on the session's stream; there is no torch in this file.
The match list of every run is compared with the reference's own list for the pair (tests/golden/config3_rows.json: sha1 of
the rows oracle/_ref produced on the full 100 Mbp pair) at every N.  The sorted-mer-list build (config 4), the gapped DP
(GCUPS, config 5) and the HMM are reported in the `sml` / `dp` / `hmm` objects of the same line; `buildindex` (N = 1) is BASELINE
config 0 end to end.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


class NvmlSampler:
    """SM clock + throttle reasons through NVML from a background thread (every ~2 ms): the timed region of this bench is tens of
    milliseconds, far shorter than nvidia-smi's sampling period.  start() returns False when NVML is unusable (the caller then
    falls back to the nvidia-smi sampler)."""

    def __init__(self, torch_device_index):
        self.idx = torch_device_index
        self.samples = []
        self._stop = False
        self._thread = None
        self._h = None
        self._nv = None

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.idx
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[self.idx])
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)  # probe
            self._h, self._nv = h, nv
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
            return True
        except Exception:  # noqa: BLE001
            return False

    def _run(self):
        nv, h = self._nv, self._h
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons(h))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop = True
        if self._thread is not None:
            self._thread.join(1.0)
        nv, h = self._nv, self._h
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            mx = None
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["no samples"], "source": "nvml"}
        bits = 0
        for _, r in self.samples:
            bits |= r
        names = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        return {"sm_mhz": float(np.median([c for c, _ in self.samples])), "sm_max_mhz": mx, "reasons": [n for n, b in names if bits & b],
                "samples": len(self.samples), "source": "nvml"}


def start_clock_sampler(gpu_index):
    s = NvmlSampler(gpu_index)
    if s.start():
        return s
    return ClockSampler(gpu_index)


def pinned_copy(lib, arr):
    from mauve_py_b200._capi import check
    p = C.c_void_p()
    check(lib.mcu_host_alloc(C.byref(p), arr.size))
    view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.size,))
    view[:] = arr
    return p, view


def workload(mbp):
    from mauve_py_b200 import synth
    n = int(mbp * 1_000_000)
    a, b = synth.config3_pair(n=n)
    return a, b


def cpu_checker():
    """the CPU arm: the reference's own code compiled in place (oracle/_ref) when present, else the C restatement"""
    import _oracle
    if _oracle.have_ref():
        return _oracle.ref_checker(), "reference"
    return _oracle.oracle_checker(), "port"


def cpu_sample(a, b, sample_bp):
    """bounded sample of the same workload for the CPU arm: the leading `sample_bp` bases of both genomes"""
    return a[:sample_bp].tobytes(), b[:sample_bp].tobytes()


def golden_config3():
    """what the reference itself returns for the full 100 Mbp pair (tests/golden/make_golden_config3.py ran oracle/_ref on it)"""
    p = os.path.join(ROOT, "tests", "golden", "config3_rows.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def run_reference_arm(args):
    """--impl reference: the reference's own match finder (oracle/_ref: unmodified sources compiled in place; else the C restatement)
    on the host cores.  Nothing of the product is imported here."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    a, b = workload(args.mbp)  # the same pair the GPU arm runs; the CPU gets its leading part
    sa, sb = cpu_sample(a, b, int(args.cpu_sample_mbp * 1e6))
    chk, kind = cpu_checker()
    weight = chk.default_seed_weight(int(args.mbp * 1e6))
    seed = chk.get_seed(weight, 3)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rows, _ = chk.find_mums(sa, sb, seed, 0)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = (len(sa) + len(sb)) / 1e6 / (ms / 1e3)
    sample = "leading %.1f Mbp of both genomes of the %g Mbp pair, %d matches; single thread (the reference has no threads)" % (
        args.cpu_sample_mbp, args.mbp, rows.shape[0])
    cfg = config_dict(args, weight, seed)
    g = golden_config3()
    cfg["reference_arm_sample"] = sample
    if g and abs(args.mbp - 100.0) < 1e-9:
        cfg["reference_arm_full_size"] = "the same code on the FULL pair: %.0f s = %.3f Mbp/s, %d rows (tests/golden/config3_rows.json)" % (
            g["reference_seconds"], (g["n0"] + g["n1"]) / 1e6 / g["reference_seconds"], g["reference"]["rows"])
    line = {
        "impl": "reference", "metric": "Mbp/s seed+match+extend", "value": value, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": 1, "kind": kind, "sample": sample, "cores_on_box": os.cpu_count(),
                         "projection_if_embarrassingly_parallel": value * (os.cpu_count() or 1),
                         "projection_note": "PROJECTION, not a measurement: the reference has no threads; value x cores_on_box"},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def config_dict(args, weight, seed):
    return {"workload": "synthetic %g Mbp pair (BASELINE config 3 / SURVEY 8d C3: 0.9%% SNP, 0.1%% indel events, inversions, translocations), "
                        "default seed weight %d rank CODING_SEED -> pattern 0x%x" % (args.mbp, weight, seed),
            "genome_bp": int(args.mbp * 1e6), "seed_weight": weight, "seed_pattern": hex(seed),
            "sharding": ("one library call per step (mcu_session_run_sharded): each rank packs 1/N of the genomes + ncclAllGather, seeds owned per rank by a "
                         "mer hash, ncclAllReduce of the 1-bit/base unique-seed bitmap, ncclSend/ncclRecv gather of match rows to rank 0, merge there")
            if args.gpus > 1 else "none",
            "l2": "inputs (2 x %g MB ASCII, %.1f GB of 8-byte seed records) exceed the 126 MB L2; no extra flush" % (args.mbp, 2 * args.mbp * 8e6 / 1e9)}


def _nw_one(p):
    import _oracle
    chk = _oracle.ref_checker() if _oracle.have_ref() else _oracle.oracle_checker()
    chk.nw_align(p[0], p[1])
    return len(p[0]) * len(p[1])


def measure_dp(mp, synth, args, rank=0, comm=None):
    """GCUPS of the gapped DP on BASELINE config 5's regions (lenA log-uniform 100 bp - 10 kbp, 5 % SNP + 1 % indel events):
    --dp-regions regions PER GPU (every rank draws its own), wall clock around mcu_nw_batch with host buffers (H2D of the sequences,
    kernels, D2H of the paths); kernel-only GCUPS beside it; the reference's NWSmall on a stratified sample of the same regions."""
    from mauve_py_b200 import dist as mdist
    world = comm.world if comm is not None else 1
    pairs = synth.dp_pairs(args.dp_regions, 100, 10000, seed=20261020 + 7919 * rank)
    arrs = synth.dp_arrays(pairs)
    mp.libmems.nw_batch_arrays(*arrs)          # warm-up: allocations
    if comm is not None:
        comm.barrier()
    t0 = time.perf_counter()
    res = mp.libmems.nw_batch_arrays(*arrs)
    wall = time.perf_counter() - t0
    cells = float(res["stats"][0])
    dev_ms = float(res["device_ms"])
    if comm is not None and world > 1:
        cells = comm.allreduce([cells], mdist.SUM)[0]
        wall, dev_ms = comm.allreduce([wall, dev_ms], mdist.MAX)
    dp = {"metric": "GCUPS gapped DP (full-matrix cells, NWSmall-exact paths)", "value": cells / wall / 1e9, "unit": "GCUPS",
          "timing": "wall clock around mcu_nw_batch with HOST buffers (H2D of the sequences, kernels, D2H of the paths), max over ranks",
          "kernel_only_gcups": cells / (dev_ms * 1e-3) / 1e9, "regions": len(pairs) * world, "regions_per_gpu": len(pairs), "cells": cells,
          "wall_ms": 1e3 * wall, "device_ms": dev_ms, "scaling": "weak" if world > 1 else None,
          "workload": "BASELINE config 5: %d regions per GPU, lenA log-uniform 100 bp-10 kbp, 5%% SNP + 1%% indel events" % len(pairs)}
    if rank != 0:
        return dp
    # DP roofline: ~12 int32 operations per cell against the INT32 issue rate MEASURED on this device by a register-resident
    # add/max microbenchmark (mcu_test_int32_peak)
    try:
        gops, pms = C.c_double(0.0), C.c_float(0.0)
        mp._capi.check(mp.lib().mcu_test_int32_peak(C.byref(gops), C.byref(pms)))
        if gops.value > 0:
            achieved = 12.0 * dp["kernel_only_gcups"] / world
            dp["roofline"] = {"bound": "int32 issue", "achieved": achieved, "peak": gops.value, "unit": "G thread-instructions/s", "frac": achieved / gops.value,
                              "ops_per_cell": 12, "peak_source": "measured: mcu_test_int32_peak (8 independent VIADDMNMX chains per thread, register "
                                                                 "resident, %.2f ms); achieved = 12 algorithmic int32 operations per cell (SURVEY.md 8d), kernel time, per GPU" % pms.value}
    except Exception as e:  # noqa: BLE001
        dp["roofline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if not args.no_cpu and world == 1:
        # stratified: the regions in length order, every k-th one
        order = sorted(range(len(pairs)), key=lambda i: len(pairs[i][0]))
        k = max(1, len(order) // max(args.dp_cpu_regions, 1))
        sample = [pairs[i] for i in order[k // 2::k]][:args.dp_cpu_regions]
        import multiprocessing as mproc
        procs = max(1, min(os.cpu_count() or 1, 16))
        t0 = time.perf_counter()
        with mproc.get_context("fork").Pool(procs) as pool:
            done = sum(pool.map(_nw_one, sample, chunksize=max(1, len(sample) // (procs * 8))))
        dtc = time.perf_counter() - t0
        _, kind = cpu_checker()
        dp["cpu_baseline"] = {"value": done / dtc / 1e9, "unit": "GCUPS", "cores": procs, "kind": kind, "per_core": done / dtc / 1e9 / procs,
                              "sample": "%d regions of the same batch, stratified by length (every %d-th in length order), one region per task over %d "
                                        "processes (the reference has no threads; regions are independent)" % (len(sample), k, procs)}
    return dp


def measure_hmm(mp, synth, args, rank=0, comm=None):
    """columns/s of the homology HMM: one column string per DP region (BASELINE config 5 scores every region); 2048 distinct strings
    tiled to --dp-regions per GPU.  Wall clock around mcu_hmm_batch with host buffers; kernel time beside it."""
    from mauve_py_b200 import dist as mdist
    world = comm.world if comm is not None else 1
    nreg = args.dp_regions
    base = min(nreg, 2048)
    rng = np.random.default_rng(20261020 + rank)
    lens = np.clip(np.exp(rng.uniform(np.log(100), np.log(10000), base)).astype(np.int64), 100, 10000)
    sym = [synth.hmm_string(int(l), seed=i + 4096 * rank, block=300) for i, l in enumerate(lens)]
    sym = (sym * ((nreg + base - 1) // base))[:nreg]
    params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
    mp.run_batch(sym, params, True)
    if comm is not None:
        comm.barrier()
    # what the reference's run() returns is the prediction string (1 B/column back); the posteriors (8 B/column) are an extra of this ABI
    t0 = time.perf_counter()
    mp.run_batch(sym, params, False)
    wall = time.perf_counter() - t0
    t0 = time.perf_counter()
    _, _, hms = mp.run_batch(sym, params, True)
    wall_post = time.perf_counter() - t0
    cols = float(sum(len(s) for s in sym))
    if comm is not None and world > 1:
        cols = comm.allreduce([cols], mdist.SUM)[0]
        wall, wall_post, hms = comm.allreduce([wall, wall_post, float(hms)], mdist.MAX)
    out = {"metric": "HomologyHMM columns/s (Forward+Backward posteriors, bfloat-faithful: bit-identical to the reference)",
           "value": cols / wall, "unit": "columns/s", "kernel_only_columns_s": cols / (hms * 1e-3), "strings": len(sym) * world, "columns": cols,
           "wall_ms": 1e3 * wall, "wall_ms_with_posteriors_back": 1e3 * wall_post, "device_ms": float(hms),
           "timing": "wall clock around mcu_hmm_batch with HOST (pageable) buffers: symbols in, the H/N prediction string back (what the reference's "
                     "run() returns); wall_ms_with_posteriors_back adds the optional posterior array (8 B/column); max over ranks"}
    if rank == 0:
        try:   # the case the aligner really has: ONE genome-sized string (LM/Islands.h:161)
            one = synth.hmm_string(args.hmm_single_columns, seed=99, block=400)
            mp.run_batch([one], params, True)
            t0 = time.perf_counter()
            _, _, ms1 = mp.run_batch([one], params, True)
            w1 = time.perf_counter() - t0
            out["single_string"] = {"columns": len(one), "wall_ms": 1e3 * w1, "device_ms": float(ms1), "columns_s": len(one) / w1}
            if not args.no_cpu and world == 1:
                chk, kind = cpu_checker()
                t0 = time.perf_counter()
                chk.hmm_run(one, params)
                out["single_string"]["cpu_columns_s"] = len(one) / (time.perf_counter() - t0)
                out["single_string"]["cpu_kind"] = kind
        except Exception as e:  # noqa: BLE001
            out["single_string"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


def measure_anchor_cols(mp, synth, args, rank=0):
    """anchor columns of alignment windows (muscle::FindAnchorColsPP): a batch of two-genome windows of 20,000 columns
    (the aligner's window size) in one mcu_anchor_cols_batch call -- one CTA per window -- and the aligner's own case, ONE window
    per call, with the CPU restatement beside both (checked equal on the spot)."""
    nwin, ncol = args.anchor_windows, 20000
    base = [synth.alignment_window(ncol, seed=900 + i + 64 * rank) for i in range(min(nwin, 64))]
    wins = [(base[i % len(base)], 1) for i in range(nwin)]
    lib = mp.lib()
    blob = np.concatenate([w[0].reshape(-1) for w in wins])
    n = len(wins)
    row_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(2 * ncol))
    col_off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(ncol))
    ncols = np.full(n, ncol, dtype=np.uint32)
    ones = np.ones(n, dtype=np.uint32)
    cols = np.zeros(n * ncol, dtype=np.uint32)
    counts = np.zeros(n, dtype=np.uint32)
    ms = C.c_float(0)

    def call(k):
        rc = lib.mcu_anchor_cols_batch(k, blob.ctypes.data, row_off.ctypes.data, ncols.ctypes.data, ones.ctypes.data, ones.ctypes.data, None, None,
                                       col_off.ctypes.data, cols.ctypes.data, counts.ctypes.data, None, None, C.byref(ms))
        if rc != 0:
            raise RuntimeError("mcu_anchor_cols_batch: %s" % lib.mcu_last_error().decode())
    call(n)
    t0 = time.perf_counter()
    call(n)
    wall = time.perf_counter() - t0
    dev = float(ms.value)
    call(1)
    t0 = time.perf_counter()
    call(1)
    wall1 = time.perf_counter() - t0
    dev1 = float(ms.value)
    ctr = np.zeros(8, dtype=np.uint64)
    lib.mcu_test_anchor_counters(ctr.ctypes.data)
    out = {"metric": "alignment columns/s through FindAnchorColsPP (per-column SP score, smoothing, best columns, merging; floats identical to the reference)",
           "windows": n, "columns_per_window": ncol, "value": n * ncol / wall, "unit": "columns/s", "wall_ms": 1e3 * wall, "device_ms": dev,
           "kernel_only_columns_s": n * ncol / (dev * 1e-3),
           "single_window": {"columns": ncol, "wall_ms": 1e3 * wall1, "device_ms": dev1,
                             "smoothing_segments_exact_chain": [int(ctr[0]), int(ctr[1])],
                             "sm_cycles": {"scoring": int(ctr[2]), "smoothing": int(ctr[3]), "best_columns": int(ctr[4]),
                                           "group_walk": int(ctr[5]) + int(ctr[6]), "picks": int(ctr[7])}},
           "timing": "wall clock around mcu_anchor_cols_batch with HOST buffers (rows in, anchor columns back); device_ms = the kernel alone"}
    if not args.no_cpu and rank == 0:
        import _oracle
        k = min(len(base), 16)
        t0 = time.perf_counter()
        want = [_oracle.anchor_cols(base[i], 1)[0] for i in range(k)]
        cpu = (time.perf_counter() - t0) / k
        call(n)
        same = all(np.array_equal(cols[i * ncol:i * ncol + int(counts[i])], want[i % len(base)]) for i in range(n) if i % len(base) < k)
        out["cpu_ms_per_window"] = 1e3 * cpu
        out["cpu_kind"] = "port (oracle/mauve_oracle.c: orc_anchor_cols, pinned on the reference's FindAnchorColsPP by tests/golden/anchor_cols.npz), 1 core"
        out["parity"] = "anchor columns identical to the CPU restatement on %d windows" % sum(1 for i in range(n) if i % len(base) < k) if same else "MISMATCH"
    return out


def measure_sml(mp, synth, args, a, pa, peak):
    """BASELINE config 4's stage at the size of this run: DNAMemorySML::Create = mcu_sml_build of genome 0 of the pair (pinned host
    sequence in, sorted list left on the device), for seed weights 11..21 at rank 0 and rank 3 (CODING_SEED).  Wall clock per call;
    pack / seed generation / radix sort from the library's CUDA events; the onesweep pass against the HBM roofline with SURVEY 8d's
    2 x (Kb + 4) bytes per pair and pass."""
    lib = mp.lib()
    out = {"metric": "Mbp/s sorted-mer-list build (DNAMemorySML::Create, one genome)", "unit": "Mbp/s", "genome_bp": int(a.size), "rows": []}
    best = None
    for w in range(11, 22, 2):
        for r in (0, 3):
            seed = mp.getSeed(w, r)
            L, wt = mp.getSeedLength(seed), mp.getSeedWeight(seed)
            n_out = C.c_uint64(0)
            mp._capi.check(lib.mcu_sml_build(pa, a.size, seed, None, None, None, C.byref(n_out)))
            t0 = time.perf_counter()
            mp._capi.check(lib.mcu_sml_build(pa, a.size, seed, None, None, None, C.byref(n_out)))
            wall = time.perf_counter() - t0
            st = np.zeros(6, dtype=np.float32)
            lib.mcu_sml_last_stats(st.ctypes.data)
            passes, kb, npos = int(st[3]), int(st[4]), float(n_out.value)
            sort_bytes = passes * 2.0 * (kb + 4) * npos
            row = {"weight_requested": w, "rank": r, "seed": hex(seed), "seed_length": L, "seed_weight": wt, "wall_ms": 1e3 * wall,
                   "mbp_s": a.size / 1e6 / wall, "device_mbp_s": a.size / 1e6 / (float(st[0] + st[1] + st[2]) * 1e-3),
                   "pack_ms": float(st[0]), "seedgen_ms": float(st[1]), "sort_ms": float(st[2]), "radix_passes": passes, "key_bytes": kb,
                   "onesweep_gbs": sort_bytes / (float(st[2]) * 1e-3) / 1e9 if st[2] > 0 else None,
                   "onesweep_frac": sort_bytes / (float(st[2]) * 1e-3) / 1e9 / peak if st[2] > 0 else None}
            out["rows"].append(row)
            if w == args.sml_headline_weight and r == 3:
                best = row
    best = best or out["rows"][-1]
    out["value"] = best["mbp_s"]
    out["headline"] = "weight %d rank 3: %.0f Mbp/s wall (H2D of the ASCII genome inside), %.0f Mbp/s device; onesweep %.0f GB/s = %.2f of measured HBM" % (
        best["weight_requested"], best["mbp_s"], best["device_mbp_s"], best["onesweep_gbs"] or 0.0, best["onesweep_frac"] or 0.0)
    return out


def measure_config4(mp, synth, args, peak):
    """BASELINE config 4 at its stated size: a synthetic 1 Gbp pair (uniform random ancestor + a 1 % SNP copy, seed 20261019).  (a) the
    sorted-mer-list build of one genome (DNAMemorySML::Create = mcu_sml_build, pageable host sequence in, sorted list left in HBM) for
    seed weights 11 / 15 / 19 / 21 at rank 0 and rank 3, with the onesweep passes against the HBM roofline (2 x (Kb + 4) bytes per pair
    and pass); (b) one seed + match + extend pass over the pair at the default weight (genomes resident)."""
    lib = mp.lib()
    n = int(args.config4_gbp * 1e9)
    rng = np.random.default_rng(20261019)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    step = 100_000_000
    a = np.concatenate([lut[rng.integers(0, 4, min(step, n - o), dtype=np.uint8)] for o in range(0, n, step)])
    out = {"metric": "Mbp/s sorted-mer-list build + seed+match+extend, BASELINE config 4", "genome_bp": n, "rows": []}
    for w in (11, 15, 19, 21):
        for r in (0, 3):
            seed = mp.getSeed(w, r)
            n_out = C.c_uint64(0)
            walls = []
            for _ in range(2):
                t0 = time.perf_counter()
                mp._capi.check(lib.mcu_sml_build(a.ctypes.data, a.size, seed, None, None, None, C.byref(n_out)))
                walls.append(time.perf_counter() - t0)
            st = np.zeros(6, dtype=np.float32)
            lib.mcu_sml_last_stats(st.ctypes.data)
            passes, kb, npos = int(st[3]), int(st[4]), float(n_out.value)
            sort_bytes = passes * 2.0 * (kb + 4) * npos
            dev_ms = float(st[0] + st[1] + st[2])
            out["rows"].append({"weight_requested": w, "rank": r, "seed": hex(seed), "seed_length": mp.getSeedLength(seed), "seed_weight": mp.getSeedWeight(seed),
                                "wall_ms": 1e3 * min(walls), "mbp_s": n / 1e6 / min(walls), "device_ms": dev_ms, "device_mbp_s": n / 1e6 / (dev_ms * 1e-3),
                                "seedgen_ms": float(st[1]), "sort_ms": float(st[2]), "radix_passes": passes, "key_bytes": kb,
                                "onesweep_gbs": sort_bytes / (float(st[2]) * 1e-3) / 1e9 if st[2] > 0 else None,
                                "onesweep_frac": sort_bytes / (float(st[2]) * 1e-3) / 1e9 / peak if st[2] > 0 else None})
    b = np.concatenate([synth.snps(a[o:o + step], 0.01, rng) for o in range(0, n, step)])
    weight = mp.getDefaultSeedWeight((a.size + b.size) // 2)
    seed = mp.getSeed(weight, mp.CODING_SEED)
    sess = mp.AnchorSession()
    sess.upload(a, b)
    sess.run(seed)
    ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        nm = sess.run(seed)
        ms.append(1e3 * (time.perf_counter() - t0))
    out["pair_step"] = {"seed_weight": weight, "seed": hex(seed), "matches": int(nm), "ms": min(ms), "device_ms": float(sess.stage_ms[6]),
                        "mbp_s": 2 * n / 1e6 / (min(ms) * 1e-3), "bucketed": bool(sess.stage_ms[7] < 0), "repeat_limit_flag": int(sess.stats[3]),
                        "parity": "size-independent properties only at this size (tests/test_zz_fullsize_gpu.py); the reference needs ~1.5 h and 40 GB for this pair"}
    sess.close()
    out["value"] = out["rows"][-3]["mbp_s"]
    return out


def measure_buildindex(mp, args):
    """BASELINE config 0 end to end: mauve.buildIndex on the MDS42 pair.  Reference arm = the reference's own progressiveMauve binary
    (oracle/_ref, unmodified sources) + the LUT construction, on the box's host cores; ours = mauve_py_b200.buildIndex (sorted mer
    lists + initial anchors on the device) driving the reference binary with the link-time seams (every gapped DP of the run, the gap
    searches of recursive anchoring on the device).  The LUT is checked against the golden one minted by the
    reference's own Python (tests/golden/mds42_lut.npz).  Wall-clock seconds; lower is better."""
    import ast
    import gzip
    import hashlib
    import shutil
    from mauve_py_b200 import buildindex as B
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    ref_bin = os.path.join(ref_dir, "progressiveMauve")
    ours_bin = next((os.path.join(ref_dir, n) for n in ("progressiveMauve_cuda_all", "progressiveMauve_cuda") if os.path.exists(os.path.join(ref_dir, n))), None)
    if not os.path.exists(ref_bin) or ours_bin is None:
        return {"unavailable": "oracle/_ref binaries not built (they need /root/reference at build time)"}
    golden = os.path.join(ROOT, "tests", "golden")
    z = np.load(os.path.join(golden, "mds42_lut.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    work = tempfile.mkdtemp()
    saved = {k: os.environ.get(k) for k in ("MAUVE_DIR", "MAUVE_CUDA_GAP_SEAM")}
    try:
        fas = []
        for name in ("mds42_recoded", "mds42_full"):
            fp = os.path.join(work, name + ".fa")
            with gzip.open(os.path.join(golden, name + ".fa.gz"), "rb") as f, open(fp, "wb") as g:
                shutil.copyfileobj(f, g)
            fas.append(fp)
        seqs = [B.getSeqFromFile(fp) for fp in fas]
        # reference arm
        t0 = time.perf_counter()
        subprocess.run([ref_bin, "--output=" + os.path.join(work, "ref.xmfa"), fas[0], fas[1]], cwd=work, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=600, check=True)
        ref_lut = B.lut_from_xmfa(os.path.join(work, "ref.xmfa"), seqs[0], seqs[1])
        t_ref = time.perf_counter() - t0
        for fp in fas:
            if os.path.exists(fp + ".sslist"):
                os.remove(fp + ".sslist")
        # ours
        bindir = os.path.join(work, "bin")
        os.makedirs(bindir)
        os.symlink(ours_bin, os.path.join(bindir, "progressiveMauveStatic"))
        os.environ["MAUVE_DIR"] = bindir
        os.environ["MAUVE_CUDA_GAP_SEAM"] = "1"
        t0 = time.perf_counter()
        lut = mp.buildIndex(fas[0], fas[1])
        t_ours = time.perf_counter() - t0
        ok = hashlib.sha1(lut.tobytes()).hexdigest() == meta["lut_sha1"] and hashlib.sha1(ref_lut.tobytes()).hexdigest() == meta["lut_sha1"]
        return {"metric": "buildIndex wall seconds, MDS42 pair (BASELINE config 0)", "unit": "s", "higher_is_better": False,
                "reference_s": t_ref, "ours_s": t_ours, "speedup": t_ref / t_ours if t_ours > 0 else None,
                "lut": "identical to the golden LUT (3,981,477 entries)" if ok else "DIFFERENT",
                "ours": "sorted mer lists + anchors on the device, %s (gapped DP, gap searches on the device)" % os.path.basename(ours_bin),
                "reference": "oracle/_ref/progressiveMauve (unmodified sources) + LUT construction, 1 thread"}
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        shutil.rmtree(work, ignore_errors=True)


_REAL_STDOUT = None


def claim_stdout():
    """Native libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.  Everything
    else goes to stderr: fd 1 is pointed at fd 2 for the run and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def spread(xs):
    xs = sorted(xs)
    return {"min": xs[0], "median": xs[len(xs) // 2], "max": xs[-1]} if xs else None


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbp", type=float, default=100.0, help="genome size of the synthetic pair in Mbp")
    ap.add_argument("--cpu-sample-mbp", type=float, default=10.0, help="leading Mbp of both genomes timed on the CPU (10-30 s of reference work)")
    ap.add_argument("--dp-regions", type=int, default=100000, help="BASELINE config 5 regions per GPU")
    ap.add_argument("--dp-cpu-regions", type=int, default=1000, help="stratified sample of the regions aligned by the reference's NWSmall")
    ap.add_argument("--hmm-single-columns", type=int, default=4000000)
    ap.add_argument("--anchor-windows", type=int, default=1184, help="windows of 20,000 columns in the FindAnchorColsPP batch (8 per SM)")
    ap.add_argument("--sml-headline-weight", type=int, default=19)
    ap.add_argument("--config4-gbp", type=float, default=1.0, help="genome size (Gbp) of the BASELINE config 4 measurement at N = 1; 0 skips it")
    ap.add_argument("--no-dp", action="store_true")
    ap.add_argument("--no-sml", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-buildindex", action="store_true", help="skip the MDS42 buildIndex end-to-end measurement (~1 minute of host time)")
    ap.add_argument("--buildindex-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.buildindex_only:   # child of the main run (see `bidx` below): prints the buildindex object alone
        import mauve_py_b200 as mp
        from mauve_py_b200._capi import check
        check(mp.lib().mcu_init(env_int("LOCAL_RANK", 0)))
        try:
            out = measure_buildindex(mp, args)
        except Exception as e:  # noqa: BLE001
            out = {"error": "%s: %s" % (type(e).__name__, e)}
        emit(out)
        return

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import mauve_py_b200 as mp
    from mauve_py_b200 import dist as mdist
    from mauve_py_b200 import synth
    from mauve_py_b200._capi import check

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    comm = mdist.init_from_env()   # mcu_init(LOCAL_RANK) + the library's own NCCL communicator
    world, rank = comm.world, comm.rank
    local = env_int("LOCAL_RANK", 0)
    if world > 1:
        args.gpus = world
    lib = mp.lib()

    a, b = workload(args.mbp)
    nbases = int(a.size + b.size)
    weight = mp.getDefaultSeedWeight((a.size + b.size) // 2)
    seed = mp.getSeed(weight, mp.CODING_SEED)
    pa, va = pinned_copy(lib, a)
    pb, vb = pinned_copy(lib, b)

    sess = mp.AnchorSession()
    sess.upload_ptr(pa, a.size, pb, b.size)   # `value`: both genomes resident in HBM (on every rank) before the timed region

    def step_resident():
        return sess.run(seed) if world == 1 else sess.run_sharded(seed)

    # ---- value: device-resident inputs --------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    launches0 = sess.launch_count()
    stage = np.zeros(16, dtype=np.float64)
    step_ms, dev_ms_steps = [], []
    sampler = start_clock_sampler(local) if rank == 0 else None   # before the barrier: its start-up (NVML init) is not part of any rank's steps
    comm.barrier()
    t0 = time.perf_counter()
    nmatch = 0
    for _ in range(args.steps):
        ts = time.perf_counter()
        nmatch = step_resident()
        step_ms.append(1e3 * (time.perf_counter() - ts))
        stage += sess.stage_ms
        dev_ms_steps.append(float(sess.stage_ms[6]))
    t_loop = time.perf_counter() - t0
    comm.barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = sess.launch_count() - launches0
    stats = sess.stats.copy()
    dt, loop_max = comm.allreduce([dt, t_loop], mdist.MAX)
    loop_min = comm.allreduce([t_loop], mdist.MIN)[0]
    rank_loop_ms = {"max": 1e3 * loop_max / args.steps, "min": 1e3 * loop_min / args.steps,
                    "closing_barrier_ms_total": 1e3 * (dt - loop_max)}   # per-step loop time of the slowest / fastest rank; what the closing barrier added to the K steps
    ms_per_step = 1e3 * dt / args.steps
    value = nbases / 1e6 / (dt / args.steps)
    stage /= args.steps

    # ---- parity of what was just timed: the merged list against the reference's own list for this pair -------------
    parity = None
    if rank == 0:
        rows = sess.download().copy()
        sha = hashlib.sha1(np.ascontiguousarray(rows, dtype=np.int64).tobytes()).hexdigest()
        g = golden_config3()
        parity = {"rows": int(rows.shape[0]), "rows_sha1": sha, "repeat_limit_flag": int(stats[3]),
                  "rows_left_out_by_the_repeat_limit_divergence": 0 if int(stats[3]) == 0 else None}
        if g and abs(args.mbp - 100.0) < 1e-9:
            parity["golden_sha1"] = g["reference"]["sha1"]
            parity["golden"] = "oracle/_ref (the reference's own MatchFinder / MemHash, unmodified sources) on the full pair: %d rows, %.0f s" % (
                g["reference"]["rows"], g["reference_seconds"])
            parity["identical"] = bool(sha == g["reference"]["sha1"])
            if not parity["identical"]:
                print("PARITY FAILURE: rows differ from the reference's list for this pair", file=sys.stderr)

    # ---- e2e: host buffers through the public C-ABI call -----------------------------------------
    cap = max(int(nmatch) + 1024, 1)
    e2e_rows = C.c_void_p()
    check(lib.mcu_host_alloc(C.byref(e2e_rows), cap * 24))
    n_out = C.c_uint64(0)

    def step_e2e():
        if world == 1:
            check(lib.mcu_find_mums_into(pa, a.size, pb, b.size, seed, 0, e2e_rows, cap, C.byref(n_out), None))
        else:
            check(lib.mcu_find_mums_sharded(pa, a.size, pb, b.size, seed, 0, e2e_rows, cap, C.byref(n_out), None))
        return int(n_out.value)

    for _ in range(2):
        step_e2e()
    e2e_ms = []
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        n_e2e = step_e2e()
        e2e_ms.append(1e3 * (time.perf_counter() - ts))
    comm.barrier()
    dte = comm.allreduce([time.perf_counter() - t0], mdist.MAX)[0]
    e2e_value = nbases / 1e6 / (dte / args.steps)
    e2e_same = None
    if rank == 0 and parity is not None:
        got = np.ctypeslib.as_array(C.cast(e2e_rows, C.POINTER(C.c_int64)), shape=(max(n_e2e, 1), 3))[:n_e2e]
        e2e_same = bool(hashlib.sha1(np.ascontiguousarray(got).tobytes()).hexdigest() == parity["rows_sha1"])
        parity["e2e_rows_identical_to_resident_rows"] = e2e_same
    def slice_bases(n):   # what one rank uploads of a genome of n bases (csrc/anchor.cuh pack_chunk_words)
        words = (n + 15) // 16 + 2
        chunk = (((words + world - 1) // world) + 3) & ~3
        return min(n, 16 * chunk)
    h2d = nbases if world == 1 else slice_bases(int(a.size)) + slice_bases(int(b.size))

    # ---- secondary metrics: every rank takes part at N > 1 (weak scaling: --dp-regions per GPU) ----
    dp = hmm = sml = anchor_cols = None
    if not args.no_dp:
        try:
            dp = measure_dp(mp, synth, args, rank, comm)
        except Exception as e:  # noqa: BLE001
            print("rank %d: DP measurement failed: %s: %s" % (rank, type(e).__name__, e), file=sys.stderr)
            dp = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            hmm = measure_hmm(mp, synth, args, rank, comm)
        except Exception as e:  # noqa: BLE001
            print("rank %d: HMM measurement failed: %s: %s" % (rank, type(e).__name__, e), file=sys.stderr)
            hmm = {"error": "%s: %s" % (type(e).__name__, e)}
        if rank == 0:
            try:
                anchor_cols = measure_anchor_cols(mp, synth, args, rank)
            except Exception as e:  # noqa: BLE001
                print("anchor column measurement failed: %s: %s" % (type(e).__name__, e), file=sys.stderr)
                anchor_cols = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- sorted mer list of genome 0, sharded by mer range over the N ranks (collective; at N = 1 the plain build) ----
    sml_sharded = None
    if not args.no_sml:
        try:
            walls, devs = [], []
            for _ in range(3):
                comm.barrier()
                t0 = time.perf_counter()
                spos, dms = mp.libmems.sml_build_sharded(a, seed)
                walls.append(time.perf_counter() - t0)
                devs.append(dms)
            w = comm.allreduce([min(walls)], mdist.MAX)[0]
            d = comm.allreduce([min(devs)], mdist.MAX)[0]
            sml_sharded = {"metric": "Mbp/s sorted-mer-list build sharded by mer range (positions gathered on rank 0)", "unit": "Mbp/s", "n_gpus": world,
                           "genome_bp": int(a.size), "seed": hex(seed), "wall_ms": 1e3 * w, "device_ms": d, "value": a.size / 1e6 / w,
                           "device_mbp_s": a.size / 1e6 / (d * 1e-3), "list_length": int(spos.size) if rank == 0 else None,
                           "note": "every rank uploads and scans the genome, sorts the seeds of its mer range, ncclSend/Recv of the 4-byte positions to "
                                   "rank 0; wall: pageable host sequence in, positions back to pageable host memory on rank 0 (400 MB: that copy is most of it); device: CUDA events from the genome being in HBM to the gathered list being in rank 0's HBM (pack, scan, sort, gather)"}
        except Exception as e:  # noqa: BLE001
            sml_sharded = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        comm.barrier()
        comm.close()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    kb = 4 if 2 * weight + 2 <= 32 else 8
    nsorted = float(stats[5])          # seed records (all ranks)
    npairs = float(stats[0])
    bucketed = sess.stage_ms[7] < 0
    traffic_tab = {}
    prof = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            traffic_tab = json.load(open(prof))
        except Exception:
            traffic_tab = {}
    other = None
    if bucketed:
        per_rank = 1.0 / world
        kern = {"bkf_scatter1_kernel": (8.0 * nsorted * per_rank + 0.25 * nbases, float(stage[9]),
                                        "level-1 partition: 2-bit genomes in, one 8-byte seed record per owned position out"),
                "bkf_scatter2_kernel": (16.0 * nsorted * per_rank, float(stage[11]), "level-2 partition of the 8-byte seed records"),
                "bk_group3_kernel": ((8.0 * nsorted + 8.0 * npairs) * per_rank, float(stage[12]),
                                     "in-bucket grouping (TMA-fed): every 8-byte seed record read once, 8 bytes per unique seed pair written")}
        dom = max(kern, key=lambda k: kern[k][1])
        bytes_per_launch, per_launch_ms, what = kern[dom]
        kname = "%s (%s)" % (dom, what)
        launches_per_step = 1
        traffic = None
        if world == 1 and abs(args.mbp - 100.0) < 1e-9:
            traffic = (traffic_tab.get("kernels", {}).get(dom) or {}).get("dram_bytes_per_launch")
        other = {}
        for name, (nbytes, ms, _w) in kern.items():
            if ms > 0:
                other[name] = {"bytes_per_launch": nbytes, "launch_ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    else:
        passes = int(sess.stage_ms[7])
        traffic = None
        kname = "rs_onesweep_kernel (one 8-bit LSD radix pass over %d key/value pairs)" % int(nsorted)
        bytes_per_launch = 2.0 * (kb + 4) * nsorted
        per_launch_ms = float(stage[2]) / max(passes, 1)
        launches_per_step = passes
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    # whole-step algorithmic bytes (LSD-sort formulation): 1.25 + B_sml(w) + (Kb+4) per base, + 28 B per seed pair
    P = (2 * weight + 1 + 7) // 8
    b_per_base = 1.25 + 0.25 + (kb + 4) + P * 2 * (kb + 4) + (kb + 4)
    step_bytes = b_per_base * nbases + 28.0 * npairs
    # bytes the bucketed formulation itself has to move: pack 1.25, records 8 written + 8 + 16 + 8 read/written, 16 + 28 per seed pair
    b_bucket = 1.25 + 0.5 + 40.0
    bucket_bytes = b_bucket * nbases + 44.0 * npairs
    dev_ms = float(stage[6])
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "launch_ms": per_launch_ms, "launches_per_step": launches_per_step,
                "note": "dominant kernel = the longest of the three partition / grouping kernels of this run (CUDA events on the launching stream); "
                        "`step` gives the whole pass against SURVEY.md 8d's algorithmic bytes AND against the bytes the bucketed pipeline itself moves" if bucketed else None,
                "other_kernels": other,
                "step": {"algorithmic_bytes": step_bytes, "bytes_per_base": b_per_base, "device_ms": dev_ms,
                         "achieved": step_bytes / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else 0.0,
                         "frac": (step_bytes / (dev_ms * 1e-3) / 1e9 / peak) if dev_ms > 0 else 0.0,
                         "frac_e2e": step_bytes / (dte / args.steps) / 1e9 / peak,
                         "bucketed_bytes_per_base": b_bucket if bucketed else None,
                         "bucketed_frac": (bucket_bytes / (dev_ms * 1e-3) / 1e9 / peak) if (bucketed and dev_ms > 0) else None,
                         # the compulsory lower bound the reference itself quotes (LM/DNAMemorySML.h:23-24): not to be conflated
                         "compulsory_bytes_per_base": 4.25,
                         "compulsory_frac": (4.25 * nbases / (dev_ms * 1e-3) / 1e9 / peak) if dev_ms > 0 else 0.0,
                         "stage_ms": {"pack": float(stage[0]), "seedgen": float(stage[1]), "sort": float(stage[2]), "join": float(stage[3]),
                                      "extend": float(stage[4]), "order": float(stage[5])},
                         "kernel_ms": {"bk_hist1": float(stage[8]), "bk_scatter1": float(stage[9]), "bk_hist2": float(stage[10]),
                                       "bk_scatter2": float(stage[11]), "bk_group": float(stage[12]), "candidate": float(stage[13]),
                                       "extend": float(stage[14])}, "spilled_records": float(stage[15])}}

    # ---- CPU baseline on a bounded sample ---------------------------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:   # rank 0 at N = 1 only (the reference arm, --impl reference, covers every N)
        try:
            chk, kind = cpu_checker()
            sa, sb = cpu_sample(a, b, int(args.cpu_sample_mbp * 1e6))
            t0 = time.perf_counter()
            rows_c, _ = chk.find_mums(sa, sb, seed, 0)
            dtc = time.perf_counter() - t0
            # parity spot check on the same sample, through the C ABI
            grows, _ = mp.libmems.find_mums(sa, sb, seed)
            same = bool(np.array_equal(grows, rows_c))
            if not same:
                print("PARITY FAILURE on the CPU sample", file=sys.stderr)
            cpu = {"value": (len(sa) + len(sb)) / 1e6 / dtc, "unit": "Mbp/s", "cores": 1, "kind": kind, "parity": "identical" if same else "FAILED",
                   "cores_on_box": os.cpu_count(), "projection_if_embarrassingly_parallel": (len(sa) + len(sb)) / 1e6 / dtc * (os.cpu_count() or 1),
                   "projection_note": "PROJECTION, not a measurement: the reference has no threads; value x cores_on_box",
                   "sample": "leading %.1f Mbp of both genomes (%d matches, GPU result %s); single thread: the reference has no threads"
                             % (args.cpu_sample_mbp, rows_c.shape[0], "identical" if same else "DIFFERENT")}
        except Exception as e:  # noqa: BLE001
            cpu = {"error": "%s: %s" % (type(e).__name__, e)}

    if world == 1 and not args.no_sml:
        try:
            sml = measure_sml(mp, synth, args, a, pa, peak)
        except Exception as e:  # noqa: BLE001
            sml = {"error": "%s: %s" % (type(e).__name__, e)}
    config4 = None
    if world == 1 and not args.no_sml and args.config4_gbp > 0:
        try:
            sess.close()
            lib.mcu_shutdown()   # every cached device buffer of this process goes back (the DP batch alone keeps tens of GB): the 1 Gbp pair takes ~80 GB
            config4 = measure_config4(mp, synth, args, peak)
        except Exception as e:  # noqa: BLE001
            config4 = {"error": "%s: %s" % (type(e).__name__, e)}
    bidx = None
    if world == 1 and not args.no_cpu and not args.no_buildindex:
        # in a child process with a time limit: it drives external binaries, and nothing there may cost the headline line
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--buildindex-only"], capture_output=True, text=True, timeout=480)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            bidx = json.loads(lines[-1]) if lines else {"error": "no result (rc %d): %s" % (r.returncode, r.stderr[-300:])}
        except Exception as e:  # noqa: BLE001
            bidx = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {
        "metric": "Mbp/s seed+match+extend", "value": value, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64" if kb == 8 else "u32",
        "data": "synthetic", "config": config_dict(args, weight, seed), "matches": int(nmatch), "seed_pairs": int(stats[0]),
        "device_ms_per_step": dev_ms, "rank_loop_ms": rank_loop_ms, "step_ms": spread(step_ms), "device_step_ms": spread(dev_ms_steps), "parity": parity,
        "timing": "value/ms_per_step: K steps between barrier + device synchronisation on both sides, max over ranks (a step has host-visible "
                  "synchronisation points of its own, so this is the whole step); step_ms: host clock per step on rank 0; device_ms_per_step, roofline "
                  "launch_ms and kernel_ms: CUDA events recorded by the library on the stream it launches on (rank 0; at N > 1 the whole step incl. "
                  "collectives and the merge)",
        "e2e": {"value": e2e_value, "unit": "Mbp/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n_e2e) * 24,
                "ms_per_step": 1e3 * dte / args.steps, "step_ms": spread(e2e_ms),
                "api": "mcu_find_mums_into(pinned host sequences -> pinned host rows): chunked H2D on a copy stream overlapped with pack + the level-1 "
                       "partition" if world == 1 else "mcu_find_mums_sharded (collective): every rank uploads 1/N of both genomes, packs it, ncclAllGather of "
                       "the packed words; rows to rank 0's pinned buffer (h2d_bytes_per_step is per rank)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "sml": sml, "sml_sharded": sml_sharded, "config4": config4, "dp": dp, "hmm": hmm, "anchor_cols": anchor_cols, "buildindex": bidx,
    }
    emit(line)
    comm.barrier()
    comm.close()
    if parity is not None and parity.get("identical") is False:
        sys.exit(3)


if __name__ == "__main__":
    main()
