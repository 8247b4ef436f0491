"""Aggregate `ncu --page source --csv --print-source cuda,sass` by CUDA source line: stall samples and executed
warp-instructions per line, per kernel.  Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_line_summary.py [top]"""
import collections
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = csv.reader(sys.stdin)
kern, hdr, agg, text = None, None, None, None
out = []


def flush():
    if not agg:
        return
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print("==", kern, "samples", tot, "warp-instr", toti)
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if len(sys.argv) > 2 and sys.argv[2] == "ins" else 0])[:top]:
        st = ", ".join("%s %.0f%%" % (k, 100.0 * c / max(v[0], 1)) for k, c in v[2].most_common(3))
        print("  line %4s  smp %5.1f%%  ins %5.1f%%  %-70s %s" % (ln, 100.0 * v[0] / tot, 100.0 * v[1] / toti, text.get(ln, "")[:70], st))


for r in rows:
    if not r:
        continue
    if r[0] in ("Kernel Name", "Function Name"):
        flush()
        kern, hdr = r[1], None
        agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
        text = {}
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or agg is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    ln = r[0]
    if not ln.strip():
        continue  # SASS rows: the line row above them already carries the totals of the line
    if r[1].strip():
        text.setdefault(ln, r[1].strip())
    try:
        s, i = int(d["# Samples"]), int(d["Instructions Executed"])
    except (ValueError, KeyError):
        continue
    a = agg[ln]
    a[0] += s
    a[1] += i
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                v = int(d[k] or 0)
            except ValueError:
                v = 0
            if v:
                a[2][k[6:]] += v
flush()
