"""GPU-side replacement for the reference's patch-cache builder.

Reference: ``scripts/precompute.py:126-239`` maps ``_sample_counts`` (one CPU
``bm4d`` call per patch, ``data_handling.py:315-335``) over a ProcessPool and
streams ``(raw, teacher, fg)`` into three memory-mapped ``.npy`` files plus
``config.json`` / ``transform.json``.  The consumers are ``CachedPatchDataset``
(``data_handling.py:1150-1190``: ``np.load(..., mmap_mode="r")`` of ``raw``,
``teacher``, ``fg``) and ``scripts/train_bm4dnet.py:14`` (required files).

:func:`write_patch_cache` keeps that on-disk contract byte for byte — same file
names, dtypes (float32 / float32 / uint8), shapes ``(N, *patch_shape)`` and the
config keys the reference's tests pin (``tests/test_precompute.py:69-81``) — but
fills ``teacher.npy`` with batched launches of the B200 denoiser instead of N
host processes.  Patch sampling (cloud reads, coherence gate, SWC masks) stays
with the caller: it hands over the uint16 patches, their per-brain offsets and,
optionally, the foreground masks.

Unlike the reference (memmaps opened ``w+``, a crash restarts from patch 0,
``precompute.py:205-213``) the writer can resume: finished patches are recorded
in ``done.npy`` and skipped on the next call.
"""
import json
import os

import numpy as np
from numpy.lib.format import open_memmap

COUNT_DTYPE = np.float32  # scripts/precompute.py `_COUNT_DTYPE`, tests/test_precompute.py:42
REQUIRED_FILES = ("raw.npy", "teacher.npy", "fg.npy", "transform.json")  # scripts/train_bm4dnet.py:14
DEFAULT_TRANSFORM_CFG = {"kind": "asinh", "params": {"offset": 0.0, "scale": 32.0}}  # precompute.py:269-272


def _default_targets(raw_u16, offsets, sigma, max_count):
    from . import api

    return api.precompute_targets(raw_u16, offsets, sigma, max_count=max_count)


def _default_fg(raw_u16, offsets):
    from . import api

    return api.make_foreground_mask(raw_u16, offsets)


def write_patch_cache(cache_dir, patches_u16, offsets, sigma_bm4d, fg=None, transform_cfg=None, split="train",
                      extra_config=None, max_count=65535.0, batch=64, resume=True, targets_fn=None, fg_fn=None):
    """Build (or finish) a patch cache in the reference's layout.

    patches_u16   (N, D, H, W) uint16 array or memmap: the sampled raw patches
    offsets       scalar or (N,) per-patch background offsets (data_handling.py:353-354)
    sigma_bm4d    noise sigma handed to BM4D (precompute.py:284: 24)
    fg            optional (N, D, H, W) foreground masks (0/1) from annotations; when absent the
                  reference's no-annotation fallback is used: make_foreground_mask(raw)
                  (data_handling.py:444, :928-929; metrics.py:32-61), computed on the GPU
    fg_fn         (raw_u16, offsets) -> mask; defaults to ``b4d.make_foreground_mask`` (tests
                  inject the CPU oracle)
    transform_cfg resolved transform cfg stamped into transform.json / config.json
    batch         patches per GPU launch
    targets_fn    (raw_u16, offsets, sigma, max_count) -> (raw f32, teacher f32); defaults
                  to the GPU path ``b4d.precompute_targets`` (tests inject the CPU oracle
                  to check the file contract without a GPU)
    Returns the number of patches computed by this call.
    """
    patches_u16 = np.asarray(patches_u16) if not isinstance(patches_u16, np.memmap) else patches_u16
    if patches_u16.ndim != 4 or patches_u16.dtype != np.uint16:
        raise ValueError("patches_u16 must be (N, D, H, W) uint16")
    n = patches_u16.shape[0]
    shape = tuple(int(s) for s in patches_u16.shape)
    off = np.broadcast_to(np.asarray(offsets, dtype=np.float32), (n,))
    if fg is not None and tuple(fg.shape) != shape:
        raise ValueError("fg must have the shape of patches_u16")
    targets_fn = targets_fn or _default_targets
    fg_fn = fg_fn or _default_fg
    tcfg = dict(transform_cfg or DEFAULT_TRANSFORM_CFG)
    os.makedirs(cache_dir, exist_ok=True)

    cfg = {
        "split": split,
        "cache_dir": cache_dir,
        "n_patches": n,
        "transform_cfg": tcfg,
        "patch_shape": list(shape[1:]),
        "sigma_bm4d": float(sigma_bm4d),
        "count_dtype": np.dtype(COUNT_DTYPE).name,
        "denoiser": "b4d (B200-native BM4D, libb4d.so)",
    }
    cfg.update(extra_config or {})
    # what a resumed run must agree on besides (N, shape, sigma): the clip, the offsets and the input itself
    # (a hash of every patch's first and last plane and of the whole of up to 64 evenly spaced patches)
    import hashlib

    hsh = hashlib.sha1()
    hsh.update(np.float64(max_count).tobytes())
    hsh.update(np.ascontiguousarray(off).tobytes())
    hsh.update(b"fg" if fg is not None else b"nofg")
    for i in range(n):
        hsh.update(np.ascontiguousarray(patches_u16[i, 0]).tobytes())
        hsh.update(np.ascontiguousarray(patches_u16[i, -1]).tobytes())
    for i in np.unique(np.linspace(0, n - 1, min(n, 64)).astype(np.int64)):
        hsh.update(np.ascontiguousarray(patches_u16[i]).tobytes())
    cfg["b4d_resume_fingerprint"] = hsh.hexdigest()
    paths = {k: os.path.join(cache_dir, k + ".npy") for k in ("raw", "teacher", "fg", "done")}
    fresh = not (resume and all(os.path.exists(p) for p in paths.values()))
    if not fresh:
        try:
            old = json.load(open(os.path.join(cache_dir, "config.json")))
            fresh = (old.get("n_patches"), old.get("patch_shape"), old.get("sigma_bm4d"),
                     old.get("b4d_resume_fingerprint")) != (n, list(shape[1:]), float(sigma_bm4d),
                                                            cfg["b4d_resume_fingerprint"])
        except Exception:
            fresh = True
    mode = "w+" if fresh else "r+"
    kw = dict(shape=shape) if fresh else {}
    raw_mm = open_memmap(paths["raw"], mode=mode, dtype=COUNT_DTYPE, **kw)
    teacher_mm = open_memmap(paths["teacher"], mode=mode, dtype=COUNT_DTYPE, **kw)
    fg_mm = open_memmap(paths["fg"], mode=mode, dtype=np.uint8, **kw)
    done_mm = open_memmap(paths["done"], mode=mode, dtype=np.uint8, **(dict(shape=(n,)) if fresh else {}))
    if fresh:
        done_mm[:] = 0
    with open(os.path.join(cache_dir, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)

    todo = np.flatnonzero(np.asarray(done_mm) == 0)
    for a in range(0, todo.size, batch):
        idx = todo[a : a + batch]
        raw, teacher = targets_fn(np.ascontiguousarray(patches_u16[idx]), off[idx], float(sigma_bm4d), max_count)
        raw_mm[idx] = np.asarray(raw, dtype=COUNT_DTYPE)
        teacher_mm[idx] = np.asarray(teacher, dtype=COUNT_DTYPE)
        if fg is None:
            fg_mm[idx] = np.asarray(fg_fn(np.ascontiguousarray(patches_u16[idx]), off[idx]), dtype=np.uint8)
        else:
            fg_mm[idx] = np.asarray(fg[idx], dtype=np.uint8)
        raw_mm.flush()
        teacher_mm.flush()
        fg_mm.flush()
        done_mm[idx] = 1  # only after the data are on disk
        done_mm.flush()
    # stamped last, as the reference does (precompute.py:236-238)
    with open(os.path.join(cache_dir, "transform.json"), "w") as f:
        json.dump(tcfg, f, indent=1)
    return int(todo.size)


def load_patch_cache(cache_dir):
    """The reader side of the contract (``CachedPatchDataset._load_cached_arrs``,
    data_handling.py:1150-1170): memory-mapped (raw, teacher, fg) + transform cfg."""
    missing = [f for f in REQUIRED_FILES if not os.path.exists(os.path.join(cache_dir, f))]
    if missing:
        raise ValueError("%s is missing required cache files: %s" % (cache_dir, missing))
    arrs = tuple(np.load(os.path.join(cache_dir, k + ".npy"), mmap_mode="r") for k in ("raw", "teacher", "fg"))
    if not (len(arrs[0]) == len(arrs[1]) == len(arrs[2]) and arrs[0].shape[1:] == arrs[1].shape[1:] == arrs[2].shape[1:]):
        raise ValueError("inconsistent cache arrays in %s" % cache_dir)
    return arrs + (json.load(open(os.path.join(cache_dir, "transform.json"))),)
