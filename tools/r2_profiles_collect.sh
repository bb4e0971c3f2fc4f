#!/bin/bash
# Turns the files a tools/r2_profile.sh session left under gpurun_out/<tag>/ into the tracked summaries under profiles/ (run here, where ncu
# reads the reports; no GPU needed).      bash tools/r2_profiles_collect.sh prof
set -u
cd "$(dirname "$0")/.."
T=gpurun_out/${1:-prof}
P=profiles
python tools/ncu_summary.py $T/partition.ncu-rep $T/finish.ncu-rep > $P/r02_ncu_enumeration_summary.txt 2>&1
python tools/ncu_summary.py $T/sml.ncu-rep > $P/r02_ncu_sml_build_summary.txt 2>&1
python tools/ncu_summary.py $T/dp.ncu-rep > $P/r02_ncu_nw_forward_summary.txt 2>&1
python tools/ncu_summary.py $T/hmm_warp.ncu-rep > $P/r02_ncu_hmm_warp_chain_summary.txt 2>&1
if [ -f $T/anchor_cols.ncu-rep ]; then
  python tools/ncu_summary.py $T/anchor_cols.ncu-rep > $P/r02_ncu_anchor_cols_summary.txt 2>&1
  echo "# phase clocks of ONE 20,000-column window per call (mcu_test_anchor_counters), three kinds of window:" >> $P/r02_ncu_anchor_cols_summary.txt
  grep -v "^==" $T/cols_phases.log >> $P/r02_ncu_anchor_cols_summary.txt
fi
grep -v "^==" $T/launches.csv > $P/r02_launches_bench_100mbp.csv
python - "$T" <<'PY'
import csv, json, sys, collections
t = sys.argv[1]
# kernel shares of the launch list (cold-cache, serialised: shares, not absolutes)
rows = [r for r in csv.reader(open("profiles/r02_launches_bench_100mbp.csv")) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    tot.setdefault(name, [0, 0.0])
    tot[name][0] += 1
    tot[name][1] += float(r[14]) / 1e6
s = sum(v[1] for v in tot.values())
with open("profiles/r02_launch_shares.txt", "w") as f:
    f.write("# kernel shares of `bench.py --steps 2 --warmup 3` (5 steps + set-up) from profiles/r02_launches_bench_100mbp.csv (ncu --metrics gpu__time_duration.sum,\n")
    f.write("# --clock-control none: per-launch times are cold-cache and serialised -- the SHARES are what to compare with the bench line's kernel_ms)\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%-60s launches %4d  total %9.3f ms  share %5.1f %%\n" % (k[:60], v[0], v[1], 100 * v[1] / s))
lines = {}
for name in ("bench_full", "bench_reference"):
    try:
        lines[name] = json.loads([l for l in open("%s/%s.json" % (t, name)).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        lines[name] = {"error": str(e)}
json.dump(lines, open("profiles/r02_bench_lines.json", "w"), indent=1)
# dram traffic of the dominant kernel for bench.py's roofline.traffic
import subprocess, io
out = subprocess.run(["ncu", "-i", "%s/partition.ncu-rep" % t, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
h = rr[0]
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
kern = {}
for r in rr[2:]:
    name = r[h.index("Kernel Name")]
    for k in ("bk_group3_kernel", "bkf_scatter1_kernel", "bkf_scatter2_kernel"):
        if k in name and k not in kern:
            rd = float(r[h.index("dram__bytes_read.sum")]) * unit[rr[1][h.index("dram__bytes_read.sum")]]
            wr = float(r[h.index("dram__bytes_write.sum")]) * unit[rr[1][h.index("dram__bytes_write.sum")]]
            kern[k] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                       "gpu_time_duration_ms_under_ncu": float(r[h.index("gpu__time_duration.sum")])}
json.dump({"capture": "%s/partition.ncu-rep (ncu --set full --clock-control none, 100 Mbp pair, one launch of each kernel)" % t,
           "summary": "profiles/r02_ncu_enumeration_summary.txt", "kernels": kern}, open("profiles/dominant_kernel_traffic.json", "w"), indent=1)
PY
ls -la $P | tail -12
