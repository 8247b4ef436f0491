"""Developer aid: print the Wiener-stage intermediates of one reference block from the debug build of the
library (make EXTRA=-DB4D_DEBUG_DUMP, libb4d_dbg.so) or from the oracle mirror, for diffing.
    python tools/dump_ref.py gpu|cpu REF > file"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
sys.path.insert(0, ROOT)
os.environ["B4D_DUMP_REF"] = sys.argv[2]
os.environ["OMP_NUM_THREADS"] = "1"
import numpy as np  # noqa: E402

rng = np.random.default_rng(7)
for shape in ((4, 4, 4), (4, 4, 7), (8, 8, 8)):
    vol = np.clip(rng.normal(300, 24, shape), 0, 65535).astype(np.uint16)
if sys.argv[1] == "gpu":
    from b4d import _lib

    _lib.LIB_PATH = os.path.join(ROOT, "aind-exaspim-image-compression_b200", "libb4d_dbg.so")
    import b4d

    dn = b4d.Denoiser(0)
    y = dn.denoise(vol, 24.0)
    import torch

    torch.cuda.synchronize()
    from oracle import np_oracle as O

    del os.environ["B4D_DUMP_REF"]
    o = O.Oracle("mirror")
    m = o.denoise(vol, 24.0)
    gn, gw = dn.debug_accumulators(vol.size)
    mn, mw = o.accumulators(vol.size)
    print("DEBUG-BUILD accumulators: numq equal", np.array_equal(gn, mn), "wmap equal", np.array_equal(gw, mw), "out equal", np.array_equal(y, m), file=sys.stderr)
else:
    from oracle import np_oracle as O

    O.Oracle("mirror").denoise(vol, 24.0)
    sys.stdout.flush()
