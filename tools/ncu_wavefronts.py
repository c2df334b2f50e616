#!/usr/bin/env python
"""Per-SASS-instruction shared-memory wavefronts / L1 tag requests / executed count from an ncu
report's source page (view=SASS), aggregated by opcode and listing the top instructions.
    python tools/ncu_wavefronts.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
items = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ex = int(r[col["Instructions Executed"]] or 0)
    wf = int(r[col["L1 Wavefronts Shared"]] or 0)
    wfi = int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    tag = int(r[col["L1 Tag Requests Global"]] or 0)
    smp = int(r[col["# Samples"]] or 0)
    a = agg[op]
    a[0] += ex; a[1] += wf; a[2] += wfi; a[3] += tag; a[4] += smp
    items.append((wf, tag, ex, smp, src))
tot_ex = sum(a[0] for a in agg.values())
print(f"total warp instructions {tot_ex}")
print(f"{'opcode':10s} {'executed':>12s} {'%':>6s} {'smem wavefronts':>16s} {'ideal':>12s} {'L1 tag req':>12s} {'samples':>8s}")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:10s} {a[0]:12d} {100*a[0]/tot_ex:6.1f} {a[1]:16d} {a[2]:12d} {a[3]:12d} {a[4]:8d}")
print("-- top shared-memory wavefront instructions")
for wf, tag, ex, smp, src in sorted(items, key=lambda t: -t[0])[:22]:
    print(f"{wf:12d} wf  {ex:10d} ex  {smp:6d} smp  {src[:90]}")
print("-- top sampled instructions (stall samples; dominant reasons)")
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ranked = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    smp = int(r[col["# Samples"]] or 0)
    reasons = sorted(((int(r[col[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:2]
    ranked.append((smp, r[col["Source"]].strip(), reasons, int(r[col["Instructions Executed"]] or 0)))
tot_s = sum(t[0] for t in ranked) or 1
for smp, src, reasons, ex in sorted(ranked, key=lambda t: -t[0])[:40]:
    print(f"{100*smp/tot_s:5.1f}%  {ex:9d} ex  {src[:70]:70s} {reasons}")
