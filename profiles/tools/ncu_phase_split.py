#!/usr/bin/env python
"""Instruction / stall-sample share per phase of step_warp_kernel from the cuda,sass source view of ncu
(same input as ncu_source_summary.py).  Line ranges follow the section comments of the kernel sources."""
import collections
import csv
import re
import sys

ROOT = __file__.rsplit("/profiles/", 1)[0]


def anchors(path, pats):
    out = {}
    for n, line in enumerate(open(path), 1):
        for k, pat in pats.items():
            if k not in out and re.search(pat, line):
                out[k] = n
    return out


W = anchors(ROOT + "/gym_dockauv_b200/csrc/dockauv_step_warp.cuh",
            {"A": r"-- phase A$", "B": r"-- phase B$", "pass1": r"---- pass 1", "pass2": r"---- pass 2",
             "clamp": r"---- clamp \(sensor", "pool": r"---- 2x2 max-pool", "C": r"-- phase C$"})
D = anchors(ROOT + "/gym_dockauv_b200/csrc/dockauv_device.cuh",
            {"sincos": r"sincos_outlined\(double x\) \{", "mth": r"^struct Mth<double>", "clip": r"T clipv\(",
             "capsule_pre": r"void capsule_pre\(", "end": r"T log_precision\("})


def classify(f, l):
    if f == "dockauv_device.cuh":
        if D["sincos"] <= l < D["sincos"] + 5:
            return "A: sincos (out of line)"
        if D["mth"] <= l < D["clip"]:
            return "libm wrappers (sqrt/div/log/atan2; A+B)"
        if l < D["capsule_pre"]:
            return "A: dynamics (RHS, RKF45, ssa)"
        if l < D["end"]:
            return "B: capsule_pre / geometry helpers"
        return "A: reward helpers"
    if f == "dockauv_env.cuh":
        return "C: reset_env + libm slow-path subroutines" if l <= 115 else "A: command filter / C: stats"
    if f == "dockauv_step_tpe.cuh":
        return "A: nav errors, obs[0:16], reward terms" if l < 186 else "C: step_finish"
    if f == "dockauv_step_warp.cuh":
        if l < W["B"]:
            return "A: pose hand-off"
        if l < W["pass1"]:
            return "B: per-kernel setup (ray table, pool map)"
        if l < W["pass2"]:
            return "B: pass 1 (obstacle pre-pass, culls, collision)"
        if l < W["clamp"]:
            return "B: pass 2 ray loop"
        if l < W["pool"]:
            return "B: clamp + OA sum"
        if l < W["C"]:
            return "B: pooling"
        return "C: finish + obs rows out"
    return f


grp, smp = collections.Counter(), collections.Counter()
cur = line = idx = None
for r in csv.reader(sys.stdin):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        idx = {n: i for i, n in enumerate(r)}
        continue
    if r[0] != "":
        line = (cur, int(r[0]))
        continue
    try:
        inst, s = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
    except Exception:
        continue
    g = classify(*line)
    grp[g] += inst
    smp[g] += s
ti, ts = sum(grp.values()), sum(smp.values())
n_envs = float(sys.argv[1]) if len(sys.argv) > 1 else 1048576.0
print(f"{'group':52s} {'inst%':>6s} {'smp%':>6s} {'warp-inst/env':>14s}")
for g, v in sorted(grp.items(), key=lambda kv: -kv[1]):
    print(f"{g:52s} {100 * v / ti:6.1f} {100 * smp[g] / ts:6.1f} {v / n_envs:14.1f}")
print(f"{'total':52s} {100.0:6.1f} {100.0:6.1f} {ti / n_envs:14.1f}")
