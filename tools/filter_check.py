"""Developer diagnostic for the filter kernels on the GPU box (not a test, not the bench):
device output against the oracle mirror per stage, with the location of the first differences."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import b4d  # noqa: E402
from b4d import synth  # noqa: E402
from oracle import np_oracle as O  # noqa: E402


def report(tag, y, m):
    eq = np.array_equal(y, m)
    d = np.abs(y.astype(np.float64) - m)
    nbad = int((y != m).sum())
    print("%-44s bit-equal %s  differing voxels %d / %d  max-abs %.4g" % (tag, eq, nbad, y.size, d.max()), flush=True)
    if not eq:
        idx = np.argwhere(y != m)[:6]
        for i in idx:
            print("    at", tuple(i), "gpu", y[tuple(i)], "mirror", m[tuple(i)])
    return eq


def accumulator_check():
    """Tiny volumes, Wiener groups of one block: numerators and weight map integer for integer."""
    rng = np.random.default_rng(7)
    for shape in ((4, 4, 4), (4, 4, 7), (8, 8, 8), (10, 9, 12)):
        vol = np.clip(rng.normal(300, 24, shape), 0, 65535).astype(np.uint16)
        for kw, okw in (({"max_stack_size_wiener": 1}, {"k_wie": 1}), ({}, {})):
            dn = b4d.Denoiser(0, b4d.BM4DProfile(**kw))
            o = O.Oracle("mirror", **okw)
            y = dn.denoise(vol, 24.0)
            m = o.denoise(vol, 24.0)
            gn, gw = dn.debug_accumulators(vol.size)
            mn, mw = o.accumulators(vol.size)
            print("acc %s %s: out equal %s, wmap equal %s (%d differ, max |d| %d), numq equal %s (%d differ, max |d| %d)" % (
                shape, kw, np.array_equal(y, m), np.array_equal(gw, mw), int((gw != mw).sum()),
                int(np.abs(gw - mw).max()), np.array_equal(gn, mn), int((gn != mn).sum()), int(np.abs(gn - mn).max())), flush=True)
            if not np.array_equal(gw, mw):
                i = np.flatnonzero(gw != mw)[:8]
                print("    wmap idx", i, "gpu", gw[i], "mirror", mw[i])
            if not np.array_equal(gn, mn):
                i = np.flatnonzero(gn != mn)[:8]
                print("    numq idx", i.tolist(), "gpu-mirror", (gn[i] - mn[i]).tolist(), "mirror", mn[i].tolist())
                import collections
                print("    histogram of gpu-mirror:", sorted(collections.Counter((gn - mn)[gn != mn].tolist()).items()))
            dn.close()


def main():
    ok = True
    accumulator_check()
    if len(sys.argv) > 1 and sys.argv[1] == "acc":
        return 0
    cases = [("const 12^3", np.full((12, 12, 12), 77, np.uint16)),
             ("synth 16x20x24", synth.vol(16, 20, 24, seed=3)),
             ("synth 21x26x31", synth.vol(21, 26, 31, seed=3)),
             ("synth 40^3", synth.vol(40, 40, 40, seed=5)),
             ("synth 64^3", synth.vol(64, 64, 64, seed=1))]
    for name, vol in cases:
        for stages in (1, 2):
            for prof, oprof in (({}, {}),
                                ({"max_stack_size_ht": 8, "max_stack_size_wiener": 16}, {"k_ht": 8, "k_wie": 16}),
                                ({"max_stack_size_ht": 32, "max_stack_size_wiener": 4}, {"k_ht": 32, "k_wie": 4}),
                                ({"max_stack_size_ht": 2, "max_stack_size_wiener": 1}, {"k_ht": 2, "k_wie": 1}),
                                ({"max_stack_size_ht": 1, "max_stack_size_wiener": 2}, {"k_ht": 1, "k_wie": 2}),
                                ({"search_window_ht": (6, 6, 6), "search_window_wiener": (7, 7, 7)},
                                 {"search_ht": 13, "search_wie": 15})):
                if prof and vol.shape[0] > 24:
                    continue
                dn = b4d.Denoiser(0, b4d.BM4DProfile(**prof), stages)
                y = dn.denoise(vol, 24.0)
                m = O.Oracle("mirror", stages=stages, **oprof).denoise(vol, 24.0)
                ok &= report("%s stages %d %s" % (name, stages, prof or ""), y, m)
                dn.close()
    vol = synth.vol(40, 40, 40, seed=5)
    raw = vol.astype(np.float32) - np.float32(36.37)
    dn = b4d.Denoiser(0)
    ok &= report("float32 offset input 40^3", dn.denoise(raw, 24.0), O.Oracle("mirror").denoise(raw, 24.0))
    x = (np.random.default_rng(1).normal(0.3, 0.05, (24, 24, 24))).astype(np.float32)
    ok &= report("float32 non-integral 24^3", dn.denoise(x, 0.05), O.Oracle("mirror").denoise(x, 0.05))
    f = O.Oracle("f64").denoise(vol, 24.0)
    y = dn.denoise(vol, 24.0)
    print("40^3 vs f64: max-abs %.4g rel-L2 %.3g" % (np.abs(y - f).max(), np.linalg.norm(y - f) / np.linalg.norm(f)))
    print("ALL BIT-EQUAL" if ok else "DIFFERENCES FOUND")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
