#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 | python ncu_source_summary.py [top_n]

Prints, per source line, warp-level instructions executed, average active threads, stall samples and the dominant
stall reasons, plus per-file totals and an SASS opcode histogram weighted by executed instructions.
"""
import collections
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = csv.reader(sys.stdin)
cur_file, hdr = None, None
line_key = None
agg = collections.defaultdict(lambda: collections.Counter())
src_text = {}
opc = collections.Counter()
opc_thr = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {n: i for i, n in enumerate(hdr)}
        stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None:
        continue
    if r[0] != "":
        line_key = (cur_file, int(r[0]))
        src_text[line_key] = r[1].strip()
        continue
    # SASS row
    try:
        inst = int(r[idx["Instructions Executed"]])
        thr = int(r[idx["Thread Instructions Executed"]])
        smp = int(r[idx["# Samples"]])
    except Exception:
        continue
    a = agg[line_key]
    a["inst"] += inst
    a["thr"] += thr
    a["samples"] += smp
    for n, i in stall_cols:
        try:
            a[n] += int(r[i])
        except Exception:
            pass
    op = r[3].strip().split()[0] if r[3].strip() else "?"
    if op.startswith("@"):
        op = r[3].strip().split()[1]
    op = op.split(".")[0]
    opc[op] += inst
    opc_thr[op] += thr

tot_inst = sum(a["inst"] for a in agg.values())
tot_smp = sum(a["samples"] for a in agg.values())
tot_thr = sum(a["thr"] for a in agg.values())
print(f"total warp instructions {tot_inst:,}  thread instructions {tot_thr:,}  avg active threads {tot_thr / max(tot_inst, 1):.1f}  samples {tot_smp:,}")
byfile = collections.Counter()
byfile_s = collections.Counter()
for (f, l), a in agg.items():
    byfile[f] += a["inst"]
    byfile_s[f] += a["samples"]
for f, v in byfile.most_common():
    print(f"  {f:28s} inst {100 * v / tot_inst:5.1f}%  samples {100 * byfile_s[f] / max(tot_smp, 1):5.1f}%")
print(f"\ntop {top_n} source lines by stall samples:")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
    stalls = sorted(((n, a[n]) for n in a if n.startswith("stall_")), key=lambda x: -x[1])[:3]
    st = " ".join(f"{n[6:]}={100 * v / max(a['samples'], 1):.0f}%" for n, v in stalls if v)
    print(f"{key[0][8:22]:14s}:{key[1]:4d} smp {100 * a['samples'] / max(tot_smp, 1):5.1f}% inst {100 * a['inst'] / tot_inst:5.1f}% "
          f"thr/inst {a['thr'] / max(a['inst'], 1):4.1f} | {st} | {src_text.get(key, '')[:70]}")
print("\nSASS opcode mix (share of executed warp instructions, avg active threads):")
for op, v in opc.most_common(22):
    print(f"  {op:10s} {100 * v / tot_inst:5.1f}%  thr/inst {opc_thr[op] / max(v, 1):4.1f}")
stall_tot = collections.Counter()
for a in agg.values():
    for n in a:
        if n.startswith("stall_"):
            stall_tot[n] += a[n]
print("\nstall reasons (share of samples):", ", ".join(f"{n[6:]}={100 * v / max(tot_smp, 1):.1f}%" for n, v in stall_tot.most_common(8)))
