"""Golden vectors for the intensity foreground mask, made by RUNNING THE REFERENCE (this container only).

    python tests/golden/make_golden_masks.py

Imports machine_learning/metrics.py straight from /root/reference/src (read-only, nothing copied) and
records make_foreground_mask(raw, k, dilate) (metrics.py:32-61) — the mask both datasets fall back to
when a patch has no annotation (data_handling.py:444, :928-929) — on raw = uint16 -> float32 - offset
(data_handling.py:353-354).  Cases: odd and even voxel counts, fractional offsets, a constant patch
(MAD 0), bright structure, k in {3, 4.5, 6}, dilate in {0, 1, 2}.  NumPy here is 2.x: Python-float
constants stay weak, so the whole statistic is float32 (NEP 50).  /root/reference does not exist on
the GPU box, so the vectors travel as tests/golden/reference_masks.npz.
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


def blob(shape, rng, amp):
    """A PSF-blurred bright tube on Gaussian background noise (in the spirit of tests/test_metrics.py)."""
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    cy, cx = shape[1] / 2.0 + 2.0 * np.sin(z / 3.0), shape[2] / 2.0
    tube = amp * np.exp(-(((y - cy) ** 2 + (x - cx) ** 2) / 6.0))
    return np.clip(np.rint(40.0 + tube + rng.normal(0.0, 12.0, shape)), 0, 65535).astype(np.uint16)


def main():
    me = _load("metrics")
    rng = np.random.default_rng(20261019)
    out = {}
    cases = [((12, 13, 11), 300.0, 0.0), ((16, 16, 16), 900.0, 37.0), ((9, 20, 14), 150.0, 36.37),
             ((24, 24, 24), 20000.0, 12.5), ((7, 5, 3), 0.0, 1.25)]
    n = 0
    for ci, (shape, amp, off) in enumerate(cases):
        u = blob(shape, rng, amp)
        if amp == 0.0:
            u[:] = 41  # constant: MAD = 0, sigma = 1.4826e-6, nothing is foreground
        out["u%d" % ci] = u
        raw = u.astype(np.float32) - off  # data_handling.py:353-354
        for k in (6.0, 3.0, 4.5):
            for dil in (0, 1, 2):
                m = me.make_foreground_mask(raw, k=k, dilate=dil)
                assert m.dtype == bool and m.shape == shape
                out["par%d" % n] = np.array([ci, off, k, dil], dtype=np.float64)
                out["m%d" % n] = np.packbits(m)
                n += 1
    out["count"] = np.array([n])
    out["numpy_version"] = np.array([np.__version__])
    np.savez_compressed(os.path.join(HERE, "reference_masks.npz"), **out)
    print("wrote %d cases" % n, "numpy", np.__version__)


if __name__ == "__main__":
    sys.exit(main())
