#!/usr/bin/env python
"""Static resource table of every kernel (registers, stack, spills, static shared memory) from the `ptxas -v` logs the build
keeps under build/obj/ (mauve_py_b200/_build.py).  No GPU needed:  python tools/ptxas_table.py > profiles/rNN_ptxas_resources.txt"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, "
                 r"(\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?"
                 r"(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?")


def demangle(sym):
    name = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", name).replace("mcu::", ""))


def main():
    logs = sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.ptxas.log")))
    if not logs:
        sys.exit("no build/obj/*.ptxas.log: run python -c 'import __graft_entry__ as g; g.build()' first")
    print("# ptxas -v resource usage of every kernel of libmauve_cuda.so (sm_100a, -O3 -lineinfo); static, no GPU (tools/ptxas_table.py)")
    print("# stack = bytes of local stack frame, spill = bytes of spill stores/loads, smem = STATIC shared memory (dynamic is set at launch)")
    print("%-12s %-60s %5s %6s %12s %8s" % ("file", "kernel", "regs", "stack", "spill st/ld", "smem"))
    n = spilled = 0
    for f in logs:
        for m in PAT.finditer(open(f).read()):
            n += 1
            spilled += int(m.group(3)) > 0 or int(m.group(4)) > 0
            print("%-12s %-60s %5d %6d %12s %8d" % (os.path.basename(f).replace(".ptxas.log", ".cu"), demangle(m.group(1))[:60], int(m.group(5)),
                                                    int(m.group(2)), "%s/%s" % (m.group(3), m.group(4)), int(m.group(8) or 0)))
    print("# %d kernels, %d with register spills" % (n, spilled))


if __name__ == "__main__":
    main()
