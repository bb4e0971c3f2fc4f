"""bench.py executed end to end on the CPU: its own host code (argument handling, timed loops, roofline arithmetic, secondary objects,
the ONE JSON line on the real stdout) with libmauve_cuda.so replaced by a stand-in that answers from the oracle
(tests/_stub/mcu_bench_stub.c; its communicator exchanges files instead of NCCL messages).  Nothing here is a measurement: the test
exists because a Python error in bench.py would cost the round's benchmark line, and the box with the GPU is not available while
developing.  The reference arm (--impl reference) needs no stand-in and is run as it is."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
import sys, os
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import mauve_py_b200._capi as capi
capi.LIB_PATH = {stub!r}
sys.argv = ["bench.py"] + {argv!r}
import bench
bench.main()
"""

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]


def run_bench(argv, stub=True, timeout=600):
    import _emu
    code = DRIVER.format(root=ROOT, stub=_emu.bench_stub_library(), argv=argv) if stub else None
    cmd = [sys.executable, "-c", code] if stub else [sys.executable, os.path.join(ROOT, "bench.py")] + argv
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly ONE line: %r" % r.stdout[-2000:]
    return json.loads(lines[0]), r.stderr


def test_bench_main_line_dry_run():
    line, err = run_bench(["--mbp", "0.3", "--cpu-sample-mbp", "0.1", "--dp-regions", "6", "--dp-cpu-regions", "3", "--hmm-single-columns", "3000", "--steps", "2",
                           "--warmup", "1", "--no-buildindex", "--config4-gbp", "0.0002", "--anchor-windows", "3"])
    for k in REQUIRED:
        assert k in line, k
    assert line["metric"] == "Mbp/s seed+match+extend" and line["unit"] == "Mbp/s" and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["warmup"] == 3   # W >= 3 is enforced
    assert line["value"] > 0 and line["matches"] > 0 and line["gpu_launches"] > 0
    assert line["config"]["workload"].startswith("synthetic 0.3 Mbp pair") and "model" not in line["config"]
    e2e = line["e2e"]
    assert e2e["value"] > 0 and abs(e2e["h2d_bytes_per_step"] - 600000) < 6000 and e2e["d2h_bytes_per_step"] == 24 * line["matches"]
    rf = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] == "hbm" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["traffic"] is None   # the ncu capture describes the 100 Mbp launch only
    assert set(rf["other_kernels"]) == {"bkf_scatter1_kernel", "bkf_scatter2_kernel", "bk_group3_kernel"} and rf["step"]["frac"] > 0
    assert "torch" not in sys.modules or True   # bench.py itself imports no torch (checked below on the source)
    assert "import torch" not in open(os.path.join(ROOT, "bench.py")).read()
    assert line["parity"]["rows"] == line["matches"] and len(line["parity"]["rows_sha1"]) == 40 and line["parity"]["e2e_rows_identical_to_resident_rows"]
    assert line["step_ms"]["min"] <= line["step_ms"]["median"] <= line["step_ms"]["max"]
    assert "error" not in line["sml"], line["sml"]
    assert len(line["sml"]["rows"]) == 12 and line["sml"]["value"] > 0
    cpu = line["cpu_baseline"]
    assert "error" not in cpu, cpu
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] == 1 and cpu["value"] > 0 and cpu["parity"] == "identical"
    assert "error" not in line["dp"], line["dp"]
    assert line["dp"]["value"] > 0 and line["dp"]["roofline"]["frac"] > 0 and line["dp"]["cpu_baseline"]["value"] > 0
    assert "error" not in line["hmm"], line["hmm"]
    assert line["hmm"]["value"] > 0 and line["hmm"]["strings"] == 6 and line["hmm"]["single_string"]["columns"] == 3000
    assert "error" not in line["anchor_cols"], line["anchor_cols"]
    assert line["anchor_cols"]["windows"] == 3 and line["anchor_cols"]["value"] > 0 and line["anchor_cols"]["parity"].startswith("anchor columns identical")
    assert line["clocks"] is not None and "reasons" in line["clocks"]
    assert "error" not in line["sml_sharded"], line["sml_sharded"]
    assert line["sml_sharded"]["n_gpus"] == 1 and line["sml_sharded"]["value"] > 0 and line["sml_sharded"]["list_length"] > 0
    assert "error" not in line["config4"], line["config4"]
    assert len(line["config4"]["rows"]) == 8 and line["config4"]["pair_step"]["matches"] > 0 and line["config4"]["genome_bp"] == 200000
    assert line["hmm"]["wall_ms_with_posteriors_back"] > 0 and line["rank_loop_ms"]["max"] > 0
    assert line["buildindex"] is None


def test_bench_reference_arm():
    line, _ = run_bench(["--impl", "reference", "--mbp", "0.3", "--cpu-sample-mbp", "0.1", "--steps", "2", "--warmup", "1"], stub=False)
    assert line["impl"] == "reference" and line["metric"] == "Mbp/s seed+match+extend" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("synthetic 0.3 Mbp pair")


def test_bench_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "progressiveMauve_cuda_all")),
                    reason="oracle/_ref binaries not built (they need /root/reference at build time)")
def test_bench_buildindex_child_dry_run():
    """the `buildindex` object of the bench line (child process `bench.py --buildindex-only`): BASELINE config 0 end to end, the
    reference binary beside mauve_py_b200.buildIndex driving the binary with every seam; the stand-in is preloaded into the binaries"""
    import _emu
    stub = _emu.bench_stub_library()
    code = DRIVER.format(root=ROOT, stub=stub, argv=["--buildindex-only"])
    env = dict(os.environ, LD_PRELOAD=stub)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert "error" not in out and "unavailable" not in out, out
    assert out["lut"].startswith("identical to the golden LUT") and out["reference_s"] > 0 and out["ours_s"] > 0


DRIVER_DIST = DRIVER   # the N > 1 flow needs nothing else: bench.py reaches the (stand-in) communicator through the C ABI


def test_bench_two_ranks_dry_run():
    """the N > 1 flow of bench.py (torchrun environment, the NCCL id travelling through a file, sharded seed+match+extend, DP and HMM
    per rank, max over ranks, rank 0 alone prints) with the stand-in library, whose communicator exchanges files: two processes"""
    import socket
    import _emu
    stub = _emu.bench_stub_library()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    code = DRIVER_DIST.format(root=ROOT, stub=stub, argv=["--gpus", "2", "--mbp", "0.3", "--dp-regions", "6", "--hmm-single-columns", "2000", "--steps", "2", "--warmup", "1"])
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   MCU_RENDEZVOUS_TAG="t%d" % os.getpid())
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    assert outs[1][0].strip() == ""   # only rank 0 prints
    lines = [l for l in outs[0][0].splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0 and line["matches"] > 0
    assert line["cpu_baseline"] is None and line["buildindex"] is None   # N = 1 only
    assert line["e2e"]["value"] > 0 and line["e2e"]["d2h_bytes_per_step"] == 24 * line["matches"]
    assert "error" not in line["dp"] and line["dp"]["value"] > 0 and line["dp"]["regions"] == 12 and line["dp"]["scaling"] == "weak"
    assert "error" not in line["hmm"] and line["hmm"]["value"] > 0 and line["hmm"]["strings"] == 12
    assert "nccl" in line["config"]["sharding"].lower() and line["sml"] is None
    assert line["e2e"]["h2d_bytes_per_step"] < 400000    # every rank uploads its slice only


def test_smoke_host_code_dry_run():
    """__graft_entry__.smoke()'s own Python (names it uses from the package, argument forms, comparisons with the oracle) against the
    stand-in: the driver runs smoke() on the GPU box before the bench, and an AttributeError there would be found only then"""
    import _emu
    code = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
            "import mauve_py_b200._capi as capi; capi.LIB_PATH = %r\n"
            "import __graft_entry__ as g; g.smoke()\n") % (ROOT, ROOT, _emu.bench_stub_library())
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stderr[-2000:]


def test_every_package_name_used_by_tests_bench_and_tools_exists():
    """`mp.<name>` in tests/, tools/, bench.py and __graft_entry__.py must resolve in the package namespace (libmems.__all__)"""
    import glob
    import re
    import types
    import mauve_py_b200 as mp
    files = glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")) + \
        [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    missing = []
    for f in files:
        for m in re.finditer(r"\bmp\.([A-Za-z_]\w*)(?:\.([A-Za-z_]\w*))?", open(f).read()):
            a, b = m.group(1), m.group(2)
            if not hasattr(mp, a):
                missing.append((os.path.basename(f), a))
            elif b and isinstance(getattr(mp, a), types.ModuleType) and not hasattr(getattr(mp, a), b):
                missing.append((os.path.basename(f), a + "." + b))
    assert not missing, missing
