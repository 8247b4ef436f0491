"""Developer aid: compare two dumps written by tools/dump_ref.py (gpu file, cpu file)."""
import struct
import sys


def load(f):
    d = {}
    for line in open(f):
        p = line.split()
        if len(p) == 5 and p[0] == "D":
            d[(int(p[1]), int(p[2]), int(p[3]))] = p[4]
    return d


def fl(h):
    return struct.unpack("f", struct.pack("I", int(h, 16)))[0]


a, b = load(sys.argv[1]), load(sys.argv[2])
names = {0: "est coef", 1: "noisy coef", 2: "shrunk", 5: "after idct", 3: "final", 4: "wsum", 6: "term", 7: "wq"}
for sec in (0, 1, 2, 5, 3, 7, 6, 4):
    keys = [k for k in b if k[0] == sec]
    miss = [k for k in keys if k not in a]
    bad = [k for k in keys if k in a and a[k] != b[k]]
    print("%-10s entries cpu %d gpu %d missing %d differing %d" % (names[sec], len(keys), len([k for k in a if k[0] == sec]), len(miss), len(bad)))
    for k in bad[:6]:
        print("     k %d v %d (z%d y%d x%d): gpu %s %.9g cpu %s %.9g" % (k[1], k[2], k[2] >> 4, (k[2] >> 2) & 3, k[2] & 3, a[k], fl(a[k]), b[k], fl(b[k])))
