"""Developer aid: random small shapes and contents through the GPU path against the CPU oracle (stage-1 match lists
and the two-stage output, bit for bit).  python tools/fuzz_shapes.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
sys.path.insert(0, ROOT)
import b4d  # noqa: E402
from oracle import np_oracle  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dn = b4d.Denoiser(0)
bad = 0
for c in range(cases):
    shape = tuple(int(x) for x in rng.integers(4, int(os.environ.get("FUZZ_MAX", "44")), 3))
    if c % 5 == 0:
        shape = shape[:2] + (int(rng.choice([8, 16, 24, 32, 40, 48])),)  # rows that satisfy the TMA stride rule
    kind = c % 4
    base = rng.integers(0, 3000)
    vol = base + rng.normal(0, 24.0, shape)
    if kind >= 1:  # a bright blob: wide / general tiles next to byte tiles
        z, y, x = (int(rng.integers(0, s)) for s in shape)
        vol[max(z - 3, 0) : z + 4, max(y - 3, 0) : y + 4, max(x - 3, 0) : x + 4] += float(rng.choice([400, 5000, 40000]))
    if kind == 3:
        vol[:, :, : shape[2] // 2] = base  # flat half: ties everywhere
    vol = np.clip(np.rint(vol), 0, 65535).astype(np.uint16)
    sigma = float(rng.choice([10.0, 24.0, 40.0]))
    o = np_oracle.Oracle("mirror")
    ok = True
    gi, gs, gc = dn.match_stage1(vol, sigma)
    oi, os_, oc = o.match_stage1(vol, sigma)
    ok &= np.array_equal(gc, oc)
    K = gi.shape[1]
    valid = np.arange(K)[None, :] < gc[:, None]
    ok &= np.array_equal(gi[valid], oi[valid]) and np.array_equal(gs[valid], os_[valid])
    y = dn.denoise(vol, sigma)
    m = o.denoise(vol, sigma)
    ok &= np.array_equal(y, m)
    if not ok:
        bad += 1
        print("MISMATCH case", c, shape, "kind", kind, "sigma", sigma, flush=True)
print("fuzz: %d cases, %d mismatches" % (cases, bad))
sys.exit(1 if bad else 0)
