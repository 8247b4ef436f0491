"""Developer smoke/diagnostic script for the GPU box (not a test, not the bench)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200")); sys.path.insert(0, ROOT)
import numpy as np
import b4d
from b4d import synth
from oracle import np_oracle as O

def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))

d = b4d.Denoiser(0)
print("pipe peaks", d.measure_pipe_peaks(), flush=True)
rng = np.random.default_rng(0)
# 1. match lists
for name, vol in [("synth 21x26x31", synth.vol(21, 26, 31, seed=3)),
                  ("noise 16^3", rng.normal(100, 24, (16, 16, 16)).clip(0, 65535).astype(np.uint16)),
                  ("const 12^3", np.full((12, 12, 12), 77, np.uint16)),
                  ("wide 20^3", rng.integers(0, 65535, (20, 20, 20)).astype(np.uint16)),
                  ("synth 40^3", synth.vol(40, 40, 40, seed=5))]:
    for sigma in (24.0,) if "wide" not in name else (24.0, 70.0):
        gi, gs, gc = d.match_stage1(vol, sigma)
        oi, os_, oc = O.Oracle("f64").match_stage1(vol, sigma)
        print("match", name, sigma, "count eq", np.array_equal(gc, oc), "idx eq", np.array_equal(gi, oi), "ssd eq", np.array_equal(gs, os_),
              "stats", d.last_match_stats(), "mean K'", gc.mean(), flush=True)
        if not np.array_equal(gi, oi):
            bad = np.nonzero((gi != oi).any(1))[0]
            print("  first bad refs", bad[:5], gi[bad[0]], oi[bad[0]], gs[bad[0]], os_[bad[0]], gc[bad[0]], oc[bad[0]])
# 2. denoise deterministic vs mirror
det = b4d.Denoiser(0, b4d.BM4DProfile(deterministic=True))
for name, vol in [("synth 21x26x31", synth.vol(21, 26, 31, seed=3)), ("synth 40^3", synth.vol(40, 40, 40, seed=5))]:
    for stages in (1, 2):
        det.set_profile(b4d.BM4DProfile(deterministic=True), stages)
        y = det.denoise(vol, 24.0)
        m = O.Oracle("mirror", stages=stages).denoise(vol, 24.0)
        f = O.Oracle("f64", stages=stages).denoise(vol, 24.0)
        print("det", name, "stages", stages, "bit-equal mirror", np.array_equal(y, m), "maxabs", np.abs(y - m).max(), "vs f64 maxabs", np.abs(y - f).max(), "relL2", rel(y, f), flush=True)
    raw = vol.astype(np.float32) - np.float32(36.37)
    det.set_profile(b4d.BM4DProfile(deterministic=True), 2)
    y = det.denoise(raw, 24.0); m = O.Oracle("mirror").denoise(raw, 24.0)
    print("det f32", name, "bit-equal", np.array_equal(y, m), "maxabs", np.abs(y - m).max(), flush=True)
# 3. fast mode vs f64
vol = synth.vol(64, 64, 64, seed=1)
y = d.denoise(vol, 24.0); f = O.Oracle("f64").denoise(vol, 24.0)
print("fast 64^3 vs f64 maxabs", np.abs(y - f).max(), "relL2", rel(y, f), "timings", d.last_timings(), flush=True)
# 4. quantize/stats
x = (rng.normal(500, 300, 100003)).astype(np.float32)
print("quant eq", np.array_equal(d.quantize(x), O.quantize_reference(x)), np.array_equal(d.quantize(x, 3.5, 37.0, 2.5), O.quantize_noise_scaled(x, 3.5, 37.0, 2.5)))
st = d.tile_stats(vol, 0.1); med, mad, sg = O.robust_sigma(vol)
print("stats", st, O.estimate_offset(vol, 0.1), med, mad, sg)
# 5. timing 128^3 x 8
big = np.stack([synth.vol(128, 128, 128, seed=1000 + i) for i in range(4)])
for rep in range(2):
    t = time.time(); yb = d.denoise(big, 24.0); dt = time.time() - t
    print("batch 4x128^3: %.3fs  %.3e vox/s" % (dt, big.size / dt), d.last_timings(), d.last_match_stats(), flush=True)
