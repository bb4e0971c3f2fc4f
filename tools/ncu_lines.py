#!/usr/bin/env python
"""Per-source-line totals of an ncu report (needs -lineinfo and --import-source on at capture time):
    python tools/ncu_lines.py report.ncu-rep kernel_name [top]
Lists 'Instructions Executed' (warp-level), stall samples and shared-memory wavefronts per CUDA source line, by samples."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    h = rows[hdr]
    c_inst, c_samp, c_wf = h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
    c_wfi = h.index("L1 Wavefronts Shared Ideal")
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= c_wfi or not r[0].strip().isdigit():   # rows with a line number carry the line's own totals
            continue
        try:
            inst, samp = int(r[c_inst]), int(r[c_samp] or 0)
            wf, wfi = int(r[c_wf] or 0), int(r[c_wfi] or 0)
        except ValueError:
            continue
        a = per.setdefault((int(r[0]), r[1].strip()[:110]), [0, 0, 0, 0])
        a[0] += inst; a[1] += samp; a[2] += wf; a[3] += wfi
    tot = sum(a[0] for a in per.values())
    tots = sum(a[1] for a in per.values())
    print("total warp instructions %d, samples %d" % (tot, tots))
    print("%8s %6s %6s %10s %10s  line" % ("inst%", "samp%", "", "smem wf", "ideal"))
    for (ln, src), a in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%7.2f%% %5.1f%% %6d %10d %10d  %d: %s" % (100.0 * a[0] / max(tot, 1), 100.0 * a[1] / max(tots, 1), a[1], a[2], a[3], ln, src))


if __name__ == "__main__":
    main()
