"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp-instructions
and stall samples per opcode, per kernel.  Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_sass_summary.py"""
import csv, sys, collections
rows = csv.reader(sys.stdin)
kern = None; hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        kern = r[1]; agg[kern] = collections.defaultdict(lambda: [0, 0, collections.Counter()]); hdr = None; continue
    if r[0] == "Address":
        hdr = r; continue
    if hdr is None or kern is None: continue
    d = dict(zip(hdr, r))
    toks = d["Source"].split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    op = toks[0].rstrip(";") if toks else "?"
    base = ".".join(op.split(".")[:2]) if op.startswith(("ATOMS", "SHFL", "LDS", "STS", "RED", "ATOMG", "F2I", "I2F", "LDG", "BAR")) else op.split(".")[0]
    a = agg[kern][base]
    a[0] += int(d["Instructions Executed"]); a[1] += int(d["# Samples"])
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = int(d[k] or 0)
            if v: a[2][k[6:]] += v
for kern, ops in agg.items():
    tot = sum(v[0] for v in ops.values()); ts = sum(v[1] for v in ops.values())
    print("==", kern, "warp-instr", tot, "samples", ts)
    for op, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:22]:
        top = ", ".join("%s %.0f%%" % (k, 100.0 * c / max(v[1], 1)) for k, c in v[2].most_common(3))
        print("  %-14s inst %5.1f%%  samples %5.1f%%   %s" % (op, 100.0 * v[0] / tot, 100.0 * v[1] / max(ts, 1), top))
