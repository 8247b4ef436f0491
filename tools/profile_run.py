"""Developer aid: a few device-resident denoise calls on the seeded benchmark volume (for ncu / A-B timing).
    python tools/profile_run.py [size] [repeats]   (B4D_LIB selects a library variant)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b4d  # noqa: E402
import bench  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
vol = bench.make_slab_device(size, 0, size, dev)
dn = b4d.Denoiser(0)
acc = {}
for i in range(reps):
    y = dn.denoise(vol, bench.SIGMA)
    tm = dn.last_timings()
    if i > 0:
        for k, (ms, nl) in tm.items():
            acc[k] = acc.get(k, 0.0) + ms / (reps - 1)
    del y
print(os.environ.get("B4D_LIB", "libb4d.so"), "size", size, {k: round(v, 3) for k, v in acc.items()}, "total %.2f ms" % sum(acc.values()))
