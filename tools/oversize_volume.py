"""One uint16 volume larger than a pass (default 1024 x 2048 x 1024 = 2 Gi voxels) through bm4d() on ONE
GPU with ordinary NumPy arrays: the library cuts it into z-slabs with halos itself.  Developer tool."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import numpy as np
import b4d
from b4d import synth

shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 2048, 1024)
small = synth.vol(128, 128, 128, seed=1000)
vol = np.ascontiguousarray(np.tile(small, tuple(s // 128 for s in shape)))
d = b4d.get_denoiser(0)
b4d.bm4d(small, 24.0)
res = {"shape": shape, "voxels": vol.size}
for rep in range(2):
    t = time.time(); y = b4d.bm4d(vol, 24.0); dt = time.time() - t
    res.setdefault("seconds", []).append(round(dt, 3))
res["voxels_per_s_host_to_host"] = vol.size / min(res["seconds"])
res["device_ms"] = {k: [round(v[0], 1), v[1]] for k, v in d.last_timings().items()}
# periodic input: every 128-plane period of the interior is identical, and equals the matching planes of a small run
ref = b4d.bm4d(np.ascontiguousarray(vol[:384, :256, :256]), 24.0)
res["interior_equals_small_run"] = bool(np.array_equal(y[128:256, 64:192, 64:192], ref[128:256, 64:192, 64:192]))
# a shift by lcm(128, 3) = 384 planes keeps both the data and the reference grid phase
res["interior_repeats_with_period_384"] = bool(np.array_equal(y[128:256, 64:192, 64:192], y[512:640, 64:192, 64:192]))
print(json.dumps(res))
