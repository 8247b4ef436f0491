"""Generate golden vectors by RUNNING THE REFERENCE (this container only).

    python tests/golden/make_golden.py

Imports machine_learning/transforms.py and machine_learning/metrics.py straight
from /root/reference/src (read-only; nothing is copied) and records their
outputs on small seeded inputs:

  quant_*    OffsetTransform(...).inverse      transforms.py:403-411  (clip + rint + uint16)
  asinh_*    AsinhTransform.inverse            transforms.py:131-152
  offset_*   estimate_offset                   transforms.py:414-438
  mask_*     make_foreground_mask(dilate=0)    metrics.py:32-61       (median / MAD sigma)

/root/reference does not exist on the GPU box, so the vectors travel as
tests/golden/reference_vectors.npz.
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/src/aind_exaspim_image_compression/machine_learning"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    tr = _load("transforms")
    me = _load("metrics")
    rng = np.random.default_rng(20261018)
    out = {}

    class Identity(tr.IntensityTransform):
        max_count = 65535.0

        def inverse_float(self, y):
            return np.asarray(y, dtype=np.float32)

    # quantize: halves (round-half-even), negatives, above range, fractional pedestals
    x = np.concatenate(
        [
            np.arange(-4, 12, dtype=np.float32) + np.float32(0.5),
            np.array([65533.5, 65534.5, 65535.0, 65535.4, 65535.5, 65536.0, 70000.0, -0.0, 0.49999997], np.float32),
            rng.normal(300.0, 400.0, 4000).astype(np.float32),
            rng.uniform(65000.0, 66000.0, 500).astype(np.float32),
        ]
    )
    out["quant_x"] = x
    offs = [0.0, 37.0, 36.37, 1.5, 100.25]
    out["quant_offsets"] = np.array(offs, dtype=np.float64)
    for i, o in enumerate(offs):
        out["quant_q%d" % i] = tr.OffsetTransform(Identity(), offset=o).inverse(x)

    a = tr.AsinhTransform(offset=0.0, scale=32.0)
    y = rng.uniform(-0.05, 1.05, 3000).astype(np.float32)
    out["asinh_y"] = y
    out["asinh_counts"] = np.asarray(a.inverse_float(y), dtype=np.float32)
    out["asinh_q"] = a.inverse(y)

    # offset percentile over non-zero voxels
    samples, pcts, vals = [], [], []
    for k in range(8):
        n = int(rng.integers(50, 6000))
        s = np.clip(rng.normal(40.0, 25.0, n), 0, 65535).astype(np.uint16)
        s[rng.random(n) < 0.15] = 0
        if k == 6:
            s[:] = 0
        if k == 7:
            s = s[:1]
        for p in (0.1, 1.0, 0.0, 50.0, 99.9):
            samples.append(s)
            pcts.append(p)
            vals.append(tr.estimate_offset(s, percentile=p))
    out["offset_n"] = np.array([len(s) for s in samples], dtype=np.int64)
    out["offset_data"] = np.concatenate(samples)
    out["offset_pct"] = np.array(pcts, dtype=np.float64)
    out["offset_val"] = np.array(vals, dtype=np.float64)

    # median / MAD sigma through the mask it thresholds
    for k in range(4):
        shp = (9 + k, 12, 11)
        raw = np.clip(rng.normal(45.0, 20.0, shp), 0, 65535).astype(np.uint16)
        raw[rng.random(shp) < 0.02] = 4000
        out["mask_raw%d" % k] = raw
        for kk in (3.0, 6.0):
            out["mask_m%d_k%d" % (k, int(kk))] = np.packbits(me.make_foreground_mask(raw, k=kk, dilate=0))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: v.shape for k, v in out.items() if k.startswith("quant_q")})


if __name__ == "__main__":
    sys.exit(main())
