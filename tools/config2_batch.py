"""BASELINE config 2 on the GPU box: 512 synthetic 128^3 uint16 patches -> BM4D training targets
(precompute path: offset subtract, denoise, clip), one B200.  Prints voxels/s and checks a few
patches bit-exactly against the CPU oracle.  Developer tool (not a test, not the bench)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import numpy as np
import b4d
from b4d import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t0 = time.time()
clean = synth.clean_tile(1000)
rng = np.random.default_rng(7)
patches = np.empty((N, 128, 128, 128), np.uint16)
for i in range(N):
    patches[i] = np.clip(np.rint(clean + rng.normal(0, 24.0, clean.shape)), 0, 65535).astype(np.uint16)
print("generated %d patches in %.1fs" % (N, time.time() - t0), flush=True)
d = b4d.get_denoiser(0)
offs = np.round(rng.uniform(30.0, 45.0, N), 2).astype(np.float32)  # per-patch background offsets
b4d.precompute_targets(patches[:4], offs[:4], 24.0)  # warm-up
t = time.time()
raw, teacher = b4d.precompute_targets(patches, offs, 24.0)
dt = time.time() - t
print(json.dumps({"config": "512 x 128^3 uint16 patches -> targets (BASELINE configs[1])", "n": N, "seconds": dt,
                  "voxels_per_s_host_to_host": patches.size / dt, "device_ms": d.last_timings(),
                  "match_stats": d.last_match_stats()}), flush=True)
assert teacher.dtype == np.float32 and teacher.shape == patches.shape and np.isfinite(teacher).all()
assert teacher.min() >= 0 and teacher.max() <= 65535
# batch == single patch, bit for bit (chunking and batching do not change results)
for i in (0, N // 2, N - 1):
    one = np.clip(b4d.bm4d(raw[i], 24.0), 0, 65535)
    assert np.array_equal(one, teacher[i]), i
print("batch == per-patch: ok")
