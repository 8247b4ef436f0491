"""Host-to-host seconds of the public calls with ordinary (pageable) NumPy arrays on the GPU box:
bm4d(volume) for one uint16 cube and precompute_targets for a patch batch.  Developer tool."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import numpy as np
import b4d
from b4d import synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
clean = synth.clean_tile(1000)
rng = np.random.default_rng(3)
reps = S // 128
small = np.clip(np.rint(clean + rng.normal(0, 24.0, clean.shape)), 0, 65535).astype(np.uint16)
vol = np.ascontiguousarray(np.tile(small, (reps, reps, reps)))
d = b4d.get_denoiser(0)
b4d.bm4d(small, 24.0)
res = {}
for rep in range(3):
    t = time.time(); y = b4d.bm4d(vol, 24.0); res.setdefault("bm4d_%d_s" % S, []).append(round(time.time() - t, 3))
    del y
res["device_ms"] = sum(v[0] for v in d.last_timings().values())
patches = vol.reshape(reps, 128, reps, 128, reps, 128).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 128, 128, 128).copy()
offs = np.round(rng.uniform(30, 45, patches.shape[0]), 2).astype(np.float32)
for rep in range(3):
    t = time.time(); r, te = b4d.precompute_targets(patches, offs, 24.0); res.setdefault("targets_s", []).append(round(time.time() - t, 3))
    del r, te
# plain memcpy speed of this host, for scale
a = np.empty(1 << 28, np.float32); t = time.time(); a[:] = 1.0; res["first_touch_GBps"] = round(a.nbytes / (time.time() - t) / 1e9, 2)
b = np.empty_like(a); b[:] = 0; t = time.time(); b[:] = a; res["memcpy_GBps"] = round(a.nbytes / (time.time() - t) / 1e9, 2)
res["cpus"] = os.cpu_count()
print(json.dumps(res))
