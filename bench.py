#!/usr/bin/env python
"""bench.py -- headline benchmark of the anchoring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mbp 100]

One "step" = one pass of seed + match + extend (pack -> two partition passes over 8-byte seed records -> in-bucket
grouping -> candidates -> extension -> reference list order; at N > 1: NCCL all-reduce of the unique-seed bitmap before
extension and NCCL gather + merge of the match rows on rank 0) over the synthetic 100 Mbp pair
(BASELINE config 3, SURVEY.md 8d "C3"; the north_star target workload, it fits one GPU).  `value`
is Mbp/s with both genomes already resident in HBM; `e2e` is the same metric through the public
C-ABI call mcu_find_mums with pinned HOST buffers (H2D + D2H inside the timed region).  N > 1 shards
the SAME pair by a seed-ownership hash ("strong" scaling).  The gapped DP (GCUPS) and the HMM are reported
in the `dp` / `hmm` objects of the same line; `buildindex` (N = 1) is BASELINE config 0 end to end: mauve.buildIndex on the MDS42
pair, the reference's own binary timed beside ours, LUT checked against the golden one.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
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
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:  # noqa: BLE001
                h = None
            if h is None:
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


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import mauve_py_b200 as mp
    a, b = workload(args.mbp)  # the same pair the GPU arm runs; the CPU gets its leading part
    sa, sb = cpu_sample(a, b, int(args.cpu_sample_mbp * 1e6))
    chk, kind = cpu_checker()
    weight = mp.getDefaultSeedWeight(int(args.mbp * 1e6))
    seed = mp.getSeed(weight, mp.CODING_SEED)
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
    line = {
        "impl": "reference", "metric": "Mbp/s seed+match+extend", "value": value, "unit": "Mbp/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config_dict(args, weight, seed),
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
            "sharding": "seed-ownership hash per rank; NCCL all-reduce of the 1-bit/base unique-seed bitmap, NCCL gather of match rows to rank 0" if args.gpus > 1 else "none",
            "l2": "inputs (2 x %g MB ASCII, %.1f GB of key/value pairs) exceed the 126 MB L2; no extra flush" % (args.mbp, 2 * args.mbp * 12e6 / 1e9)}


def measure_dp(mp, synth, args):
    """GCUPS of the gapped DP on a sample of BASELINE config 5 (+ the reference's NWSmall on a few regions of the same batch)"""
    pairs = synth.dp_pairs(args.dp_regions, 100, 10000, seed=20261020)
    arrs = synth.dp_arrays(pairs)
    mp.libmems.nw_batch_arrays(*arrs)
    res = mp.libmems.nw_batch_arrays(*arrs)
    cells = float(res["stats"][0])
    dp = {"metric": "GCUPS gapped DP (full-matrix cells, NWSmall-exact paths)", "value": cells / (res["device_ms"] * 1e-3) / 1e9,
          "unit": "GCUPS", "regions": len(pairs), "cells": cells, "device_ms": res["device_ms"],
          "workload": "BASELINE config 5 sample: %d regions, lenA log-uniform 100 bp-10 kbp, 5%% SNP + 1%% indel events" % len(pairs)}
    # DP roofline: ~12 int32 operations per cell against the INT32 issue rate MEASURED on this device by a register-resident
    # add/max microbenchmark (mcu_test_int32_peak); the kernel's own count is ~22 instructions per cell
    try:
        gops, pms = C.c_double(0.0), C.c_float(0.0)
        mp._capi.check(mp.lib().mcu_test_int32_peak(C.byref(gops), C.byref(pms)))
        if gops.value > 0:
            achieved = 12.0 * dp["value"]
            dp["roofline"] = {"bound": "int32 issue", "achieved": achieved, "peak": gops.value, "unit": "G thread-instructions/s", "frac": achieved / gops.value,
                              "ops_per_cell": 12, "peak_source": "measured: mcu_test_int32_peak (8 independent VIADDMNMX chains per thread, register "
                                                                 "resident, %.2f ms); achieved = 12 algorithmic int32 operations per cell (SURVEY.md 8d)" % pms.value}
    except Exception as e:  # noqa: BLE001
        dp["roofline"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if not args.no_cpu:
        chk, kind = cpu_checker()
        small = [p for p in pairs if len(p[0]) <= 3000][:12]
        t0 = time.perf_counter()
        for x, y in small:
            chk.nw_align(x, y)
        dtc = time.perf_counter() - t0
        dp["cpu_baseline"] = {"value": sum(len(x) * len(y) for x, y in small) / dtc / 1e9, "unit": "GCUPS", "cores": 1, "kind": kind,
                              "sample": "%d regions <= 3 kbp of the same batch" % len(small)}
    return dp


def measure_hmm(mp, synth, args):
    """columns/s of the homology HMM: one column string per region (BASELINE config 5 scores every region), 512 distinct strings
    tiled to 32768"""
    pairs = synth.dp_pairs(512, 100, 10000, seed=20261020)
    sym = [synth.hmm_string(len(p[0]), seed=i, block=300) for i, p in enumerate(pairs)] * 64
    params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
    mp.run_batch(sym, params, True)
    _, _, hms = mp.run_batch(sym, params, True)
    return {"metric": "HomologyHMM columns/s (Forward+Backward posteriors, bfloat-faithful: bit-identical to the reference)",
            "value": sum(len(s) for s in sym) / (hms * 1e-3), "unit": "columns/s", "strings": len(sym), "device_ms": hms}


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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mbp", type=float, default=100.0, help="genome size of the synthetic pair in Mbp")
    ap.add_argument("--cpu-sample-mbp", type=float, default=10.0, help="leading Mbp of both genomes timed on the CPU (10-30 s of reference work)")
    ap.add_argument("--dp-regions", type=int, default=1536)
    ap.add_argument("--no-dp", action="store_true")
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

    import torch
    import torch.distributed as dist
    import mauve_py_b200 as mp
    from mauve_py_b200 import dist as mdist
    from mauve_py_b200 import synth
    from mauve_py_b200._capi import check

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    lib = mp.lib()
    check(lib.mcu_init(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a, b = workload(args.mbp)
    nbases = int(a.size + b.size)
    weight = mp.getDefaultSeedWeight((a.size + b.size) // 2)
    seed = mp.getSeed(weight, mp.CODING_SEED)
    L = mp.getSeedLength(seed)
    pa, va = pinned_copy(lib, a)
    pb, vb = pinned_copy(lib, b)

    sess = mp.AnchorSession()
    sess.upload_ptr(pa, a.size, pb, b.size)
    shard, nshard = mdist.shard_of(rank, world)

    def step_resident():
        if world == 1:
            return sess.run(seed, shard, nshard), None
        return mdist.run_sharded(sess, seed, rank, world), None

    # ---- value: device-resident inputs --------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    launches0 = sess.launch_count()
    stage = np.zeros(16, dtype=np.float64)
    barrier()
    sampler = start_clock_sampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    nmatch = 0
    for _ in range(args.steps):
        nmatch, _m = step_resident()
        stage += sess.stage_ms
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    launches = sess.launch_count() - launches0
    stats = sess.stats.copy()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    ms_per_step = 1e3 * dt / args.steps
    value = nbases / 1e6 / (dt / args.steps)
    stage /= args.steps

    # ---- e2e: host buffers through the public C-ABI call -----------------------------------------
    def step_e2e():
        if world == 1:
            out = C.POINTER(mp._capi.Match)()
            n_out = C.c_uint64(0)
            check(lib.mcu_find_mums(pa, a.size, pb, b.size, seed, 0, C.byref(out), C.byref(n_out), None))
            n = n_out.value
            lib.mcu_free(out)
            return n, n * 24
        sess.upload_ptr(pa, a.size, pb, b.size)
        n, merged = step_resident()
        if rank == 0 and n:   # the merged list is the step's result: read it back into pinned host memory
            assert n <= e2e_rows_cap, "match count changed between steps"
            sess.download_ptr(e2e_rows)
        return n, (n * 24 if rank == 0 else 0)

    e2e_rows, e2e_rows_cap = None, 0
    if world > 1 and rank == 0:
        e2e_rows_cap = max(int(nmatch), 1)
        e2e_rows = C.c_void_p()
        check(lib.mcu_host_alloc(C.byref(e2e_rows), e2e_rows_cap * 24))

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        n_e2e, d2h = step_e2e()
    barrier()
    dte = time.perf_counter() - t0
    t = torch.tensor([dte], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dte = float(t.item())
    e2e_value = nbases / 1e6 / (dte / args.steps)

    # ---- N > 1: gapped DP and HMM divided among the ranks (independent regions / strings, LPT; no collective on the data path) ----
    # every rank reaches the two all-reduces below whatever happened in its own share, so a failure cannot leave the others waiting
    dp_sharded = hmm_sharded = None
    if world > 1 and not args.no_dp:
        vals = [0.0, 0.0, 0.0, 0.0, 1.0]   # dp cells, hmm columns | dp ms, hmm ms | ok
        try:
            pairs = synth.dp_pairs(args.dp_regions, 100, 10000, seed=20261020)
            mine = [pairs[i] for i in mdist.lpt_partition([len(x) * len(y) for x, y in pairs], world)[rank]]
            if mine:
                arrs = synth.dp_arrays(mine)
                mp.libmems.nw_batch_arrays(*arrs)
                res = mp.libmems.nw_batch_arrays(*arrs)
                vals[0], vals[2] = float(res["stats"][0]), float(res["device_ms"])
            sym = [synth.hmm_string(len(p_[0]), seed=i, block=300) for i, p_ in enumerate(synth.dp_pairs(512, 100, 10000, seed=20261020))] * 64
            smine = [sym[i] for i in mdist.lpt_partition([len(x) for x in sym], world)[rank]]
            if smine:
                params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
                mp.run_batch(smine, params, True)
                _p, _q, hms = mp.run_batch(smine, params, True)
                vals[1], vals[3] = float(sum(len(x) for x in smine)), float(hms)
        except Exception as e:  # noqa: BLE001
            print("rank %d: sharded DP/HMM measurement failed: %s: %s" % (rank, type(e).__name__, e), file=sys.stderr)
            vals = [0.0, 0.0, 0.0, 0.0, 0.0]
        tsum = torch.tensor(vals[:2], dtype=torch.float64, device="cuda")
        tmax = torch.tensor([vals[2], vals[3], -vals[4]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum, tmax = tsum.tolist(), tmax.tolist()
        if tmax[2] == -1.0 and tmax[0] > 0 and tmax[1] > 0:   # every rank succeeded
            dp_sharded = {"metric": "GCUPS gapped DP (full-matrix cells, NWSmall-exact paths)", "value": tsum[0] / (tmax[0] * 1e-3) / 1e9, "unit": "GCUPS",
                          "cells": tsum[0], "device_ms": tmax[0], "sharding": "regions divided among %d ranks by LPT on lenA*lenB; device ms = max over ranks" % world}
            hmm_sharded = {"metric": "HomologyHMM columns/s (bfloat-faithful)", "value": tsum[1] / (tmax[1] * 1e-3), "unit": "columns/s",
                           "device_ms": tmax[1], "sharding": "strings divided among %d ranks by LPT on length; device ms = max over ranks" % world}
        else:
            dp_sharded = hmm_sharded = {"error": "a rank failed (see stderr)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = measured_peaks()
    kb = 4 if 2 * weight + 2 <= 32 else 8
    nsorted = int(stats[5])
    bucketed = sess.stage_ms[7] < 0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    other = None
    if bucketed:
        # dominant kernel of the step by time: bk_group_kernel reads every 8-byte seed record once and writes 8 bytes per unique seed pair
        npairs = float(stats[0])
        kname = "bk_group_kernel (in-bucket grouping of %d 8-byte seed records into %d unique seed pairs)" % (nsorted, int(npairs))
        bytes_per_launch = 8.0 * nsorted + 8.0 * npairs
        per_launch_ms = float(stage[12])
        launches_per_step = 1
        if world > 1 or abs(args.mbp - 100.0) > 1e-9:
            traffic = None  # the ncu capture under profiles/ is of the unsharded 100 Mbp launch
        # the two partition passes, same formula (algorithmic bytes / CUDA-event time)
        other = {}
        for name, nbytes, ms in (("bkf_scatter1_kernel", 8.0 * nsorted + 0.25 * nbases, float(stage[9])),
                                 ("bkf_scatter2_kernel", 16.0 * nsorted, float(stage[11]))):
            if ms > 0:
                other[name] = {"bytes_per_launch": nbytes, "launch_ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}
    else:
        passes = int(sess.stage_ms[7])
        traffic = None  # the capture under profiles/ describes bk_group
        kname = "rs_onesweep_kernel (one 8-bit LSD radix pass over %d key/value pairs)" % nsorted
        bytes_per_launch = 2.0 * (kb + 4) * nsorted
        per_launch_ms = float(stage[2]) / max(passes, 1)
        launches_per_step = passes
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    # whole-step algorithmic bytes (LSD-sort formulation): 1.25 + B_sml(w) + (Kb+4) per base, + 28 B per seed pair
    P = (2 * weight + 1 + 7) // 8
    b_per_base = 1.25 + 0.25 + (kb + 4) + P * 2 * (kb + 4) + (kb + 4)
    step_bytes = b_per_base * nbases + 28.0 * float(stats[0])
    # bytes the bucketed formulation actually has to move: pack 1.25, records 8 written + 8 + 16 + 8 read/written, 16 + 28 per seed pair
    b_bucket = 1.25 + 0.5 + 40.0
    dev_ms = float(stage[6])
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "launch_ms": per_launch_ms, "launches_per_step": launches_per_step,
                "note": "bk_group is issue-bound (ncu: 81 % of issue slots busy, profiles/r01_ncu_bucket_enumeration_summary.txt), so its HBM fraction is low by construction; "
                        "`step` gives the whole pass against SURVEY.md 8d's algorithmic bytes" if bucketed else None,
                "other_kernels": other,
                "step": {"algorithmic_bytes": step_bytes, "bytes_per_base": b_per_base, "device_ms": dev_ms,
                         "achieved": step_bytes / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else 0.0,
                         "frac": (step_bytes / (dev_ms * 1e-3) / 1e9 / peak) if dev_ms > 0 else 0.0,
                         "bucketed_bytes_per_base": b_bucket if bucketed else None,
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
            rows, _ = chk.find_mums(sa, sb, seed, 0)
            dtc = time.perf_counter() - t0
            # parity spot check on the same sample, through the C ABI
            grows, _ = mp.libmems.find_mums(sa, sb, seed)
            same = bool(np.array_equal(grows, rows))
            if not same:
                print("PARITY FAILURE on the CPU sample", file=sys.stderr)
            cpu = {"value": (len(sa) + len(sb)) / 1e6 / dtc, "unit": "Mbp/s", "cores": 1, "kind": kind, "parity": "identical" if same else "FAILED",
                   "cores_on_box": os.cpu_count(), "projection_if_embarrassingly_parallel": (len(sa) + len(sb)) / 1e6 / dtc * (os.cpu_count() or 1),
                   "projection_note": "PROJECTION, not a measurement: the reference has no threads; value x cores_on_box",
                   "sample": "leading %.1f Mbp of both genomes (%d matches, GPU result %s); single thread: the reference has no threads"
                             % (args.cpu_sample_mbp, rows.shape[0], "identical" if same else "DIFFERENT")}
        except Exception as e:  # noqa: BLE001
            cpu = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- gapped DP + HMM (secondary metrics of BASELINE.json) ----------------------------------------
    dp = hmm = None
    if world > 1:
        dp, hmm = dp_sharded, hmm_sharded
    elif not args.no_dp:
        # secondary metrics must never cost the headline line: a failure is reported inside the object
        try:
            dp = measure_dp(mp, synth, args)
        except Exception as e:  # noqa: BLE001
            dp = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            hmm = measure_hmm(mp, synth, args)
        except Exception as e:  # noqa: BLE001
            hmm = {"error": "%s: %s" % (type(e).__name__, e)}
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
        "device_ms_per_step": dev_ms,
        "timing": "value/ms_per_step: K steps between barrier + device synchronisation on both sides, max over ranks (a step has host-visible "
                  "synchronisation points of its own, so this is the whole step); device_ms_per_step, roofline launch_ms and kernel_ms: CUDA "
                  "events recorded by the library on the stream it launches on (rank 0)",
        "e2e": {"value": e2e_value, "unit": "Mbp/s", "h2d_bytes_per_step": int(nbases), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * dte / args.steps, "api": "mcu_find_mums(host buffers)" if world == 1 else "mcu_session_upload + enumerate, NCCL all-reduce of the seed bitmap, finish, NCCL gather, mcu_session_merge, mcu_session_download (rank 0)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "dp": dp, "hmm": hmm, "buildindex": bidx,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
