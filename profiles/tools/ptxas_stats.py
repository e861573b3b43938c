#!/usr/bin/env python
"""Compile one .cu for sm_100a with -Xptxas -v and print registers / stack / spills per kernel (no GPU needed).

    python profiles/tools/ptxas_stats.py gym_dockauv_b200/csrc/dockauv_kernels_f64.cu [-DNAME=VALUE ...] [--filter regex]
"""
import re
import subprocess
import sys


def main():
    args = sys.argv[1:]
    flt = None
    if "--filter" in args:
        k = args.index("--filter")
        flt = re.compile(args[k + 1])
        del args[k:k + 2]
    src, extra = args[0], args[1:]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "--expt-relaxed-constexpr", "-Xptxas=-v", *extra, "-c", src, "-o", "/dev/null"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stderr)
    name = None
    stack = ""
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(dockauv::KParams<.*", "", name).replace("dockauv::", "").replace("void ", "")
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack = f"stack {m.group(1):>4} spill st {m.group(2):>4} ld {m.group(3):>4}"
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            sm = re.search(r"(\d+) bytes smem", line)
            if flt is None or flt.search(name):
                print(f"{name:60s} regs {m.group(1):>3}  {stack}  smem {sm.group(1) if sm else 0}")
            name = None


if __name__ == "__main__":
    main()
