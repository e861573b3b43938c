#!/usr/bin/env python
"""Per-phase instruction / stall-sample shares of one kernel from `ncu --page source --csv --print-source cuda,sass`,
with every SASS address counted ONCE (the source page lists an inlined instruction under each level of its inline
stack) and attributed to its outermost line in dockauv_step_warp.cuh when it has one.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 | \
        python ncu_dedup_split.py [n_envs] [top_n]
"""
import collections
import csv
import sys

n_envs = float(sys.argv[1]) if len(sys.argv) > 1 else 1048576.0
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ROOT = __file__.rsplit("/profiles/", 1)[0]
cur = line = idx = None
occ = collections.defaultdict(list)
inst, smp, thr, sass = {}, {}, {}, {}
stalls = collections.defaultdict(collections.Counter)
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
        scols = [(n, i) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if r[0] != "":
        line = (cur, int(r[0]))
        continue
    a = r[2]
    try:
        inst[a] = int(r[idx["Instructions Executed"]])
        smp[a] = int(r[idx["# Samples"]])
        thr[a] = int(r[idx["Thread Instructions Executed"]])
    except Exception:
        continue
    sass[a] = r[3].split()[0] if r[3].split() else "?"
    if sass[a].startswith("@"):
        sass[a] = r[3].split()[1]
    occ[a].append(line)
    for n, i in scols:
        try:
            stalls[a][n] = int(r[i])
        except Exception:
            pass

src = {}
for f in ("dockauv_step_warp.cuh", "dockauv_step_pipe.cuh", "dockauv_step_tpe.cuh", "dockauv_device.cuh", "dockauv_env.cuh"):
    try:
        src[f] = open(f"{ROOT}/gym_dockauv_b200/csrc/{f}").read().split("\n")
    except OSError:
        src[f] = []


def owner(lines):
    for f in ("dockauv_step_pipe.cuh", "dockauv_step_warp.cuh"):
        w = [l for l in lines if l[0] == f]
        if w:
            return w[-1]
    t = [l for l in lines if l[0] == "dockauv_step_tpe.cuh"]
    if t:
        return t[-1]
    return lines[-1]


by_line = collections.Counter()
by_line_smp = collections.Counter()
by_line_thr = collections.Counter()
by_line_st = collections.defaultdict(collections.Counter)
opc = collections.Counter()
for a, ls in occ.items():
    o = owner(ls)
    by_line[o] += inst[a]
    by_line_smp[o] += smp[a]
    by_line_thr[o] += thr[a]
    by_line_st[o].update(stalls[a])
    opc[sass[a].split(".")[0]] += inst[a]
ti, ts, tt = sum(inst.values()), sum(smp.values()), sum(thr.values())
print(f"warp instructions {ti:,} ({ti / n_envs:.1f} per env)  avg active threads {tt / max(ti, 1):.1f}  samples {ts:,}")
print(f"\ntop {top_n} owner lines by instructions:")
for (f, l), v in by_line.most_common(top_n):
    st = by_line_st[(f, l)]
    tot = sum(st.values()) or 1
    top = " ".join(f"{k[6:]}={100 * c / tot:.0f}%" for k, c in st.most_common(3))
    text = src.get(f, [])
    text = text[l - 1].strip()[:70] if 0 < l <= len(text) else ""
    print(f"{f[8:-4]:10s}:{l:4d} inst {100 * v / ti:5.1f}% ({v / n_envs:6.1f}/env) smp {100 * by_line_smp[(f, l)] / ts:5.1f}% "
          f"thr {by_line_thr[(f, l)] / max(v, 1):4.1f} | {top} | {text}")
print("\nopcodes:", "  ".join(f"{k} {100 * v / ti:.1f}%" for k, v in opc.most_common(24)))
agg = collections.Counter()
for a in stalls:
    agg.update(stalls[a])
tot = sum(agg.values()) or 1
print("stalls:", "  ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in agg.most_common(9)))

# ---- phase shares: owner lines in dockauv_step_warp.cuh grouped by the section comments of the kernel
import re
marks = []
for n, text in enumerate(src["dockauv_step_warp.cuh"], 1):
    m = re.search(r"// -+ (phase [ABC])$|// ---- (pass 1|pass 2|clamp|2x2 max-pool|rows of envs)", text)
    if m:
        marks.append((n, (m.group(1) or m.group(2))))
    if "T my_oa_dot = p.sum_beta_oa" in text:
        marks.append((n, "B setup (ray table, pool map, first obstacle load)"))
    if "for (int eb = 0; eb < n_warp; eb += epp)" in text:
        marks.append((n, "B sub-batch loop head"))
marks.sort()
grp, gsmp = collections.Counter(), collections.Counter()
for (f, l), v in by_line.items():
    if f == "dockauv_step_warp.cuh":
        name = "kernel entry"
        for n, g in marks:
            if l >= n:
                name = g
    else:
        name = f
    grp[name] += v
    gsmp[name] += by_line_smp[(f, l)]
print("\nphase shares (deduplicated):")
for g, v in grp.most_common():
    print(f"  {g:55s} inst {100 * v / ti:5.1f}% ({v / n_envs:6.1f}/env)  smp {100 * gsmp[g] / ts:5.1f}%")
