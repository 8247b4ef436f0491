"""Per-call latency of bm4d() on single small patches (BASELINE configs[0]: one 64^3 patch), NumPy in/out,
float32 offset-subtracted input as the precompute path passes it.  Developer tool."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import numpy as np
import b4d
from b4d import synth
res = {}
d = b4d.get_denoiser(0)
for side in (64, 128):
    u = synth.vol(side, side, side, seed=1)
    raw = u.astype(np.float32) - np.float32(37.0)
    for name, x in (("u16", u), ("f32", raw)):
        b4d.bm4d(x, 24.0)
        ts = []
        for _ in range(30):
            t = time.perf_counter(); b4d.bm4d(x, 24.0); ts.append(time.perf_counter() - t)
        ts.sort()
        res["%d^3_%s_ms_median" % (side, name)] = round(ts[len(ts) // 2] * 1e3, 3)
        res["%d^3_%s_device_ms" % (side, name)] = round(sum(v[0] for v in d.last_timings().values()), 3)
        res["%d^3_%s_launches" % (side, name)] = sum(v[1] for v in d.last_timings().values())
print(json.dumps(res))
