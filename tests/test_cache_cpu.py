"""The patch-cache file contract — scripts/precompute.py:32-38, :170-238 on
the writer side, CachedPatchDataset (data_handling.py:1150-1190) and
scripts/train_bm4dnet.py:14 on the reader side.  The CPU oracle is injected as the target
function here (test infrastructure); the product default is the GPU path."""
import json
import os

import numpy as np
import pytest


def _oracle_targets(oracle_lib):
    def fn(raw_u16, offsets, sigma, max_count):
        raw = raw_u16.astype(np.float32) - np.asarray(offsets, np.float32)[:, None, None, None]
        o = oracle_lib.Oracle("mirror")
        teacher = np.stack([np.clip(o.denoise(r, sigma), 0, max_count) for r in raw])
        return raw, teacher.astype(np.float32)

    return fn


def _oracle_fg(oracle_lib):
    def fn(raw_u16, offsets):
        off = np.asarray(offsets, np.float32)
        return np.stack([oracle_lib.make_foreground_mask_reference(oracle_lib.read_counts(r, o)) for r, o in zip(raw_u16, off)])

    return fn


def test_cache_layout_matches_reference_contract(tmp_path, oracle_lib):
    from b4d import cache, synth

    patches = np.stack([synth.vol(8, 9, 10, seed=s) for s in range(5)])
    offs = np.array([37.0, 37.0, 12.5, 0.0, 36.37], np.float32)
    fg = (patches > 60).astype(np.uint8)
    d = str(tmp_path / "train")
    n = cache.write_patch_cache(d, patches, offs, 24.0, fg=fg, targets_fn=_oracle_targets(oracle_lib), batch=2)
    assert n == 5
    for f in ("raw.npy", "teacher.npy", "fg.npy", "transform.json", "config.json"):  # train_bm4dnet.py:14
        assert os.path.exists(os.path.join(d, f))
    raw, teacher, fgm, tcfg = cache.load_patch_cache(d)
    assert raw.dtype == np.float32 and teacher.dtype == np.float32 and fgm.dtype == np.uint8  # precompute.py:204-213
    assert raw.shape == teacher.shape == fgm.shape == (5, 8, 9, 10)
    assert np.array_equal(raw, patches.astype(np.float32) - offs[:, None, None, None])  # data_handling.py:353-354
    assert teacher.min() >= 0.0 and teacher.max() <= 65535.0  # data_handling.py:333
    assert np.array_equal(fgm, fg)
    cfg = json.load(open(os.path.join(d, "config.json")))
    assert cfg["sigma_bm4d"] == 24.0 and cfg["count_dtype"] == "float32" and cfg["patch_shape"] == [8, 9, 10]
    assert cfg["transform_cfg"] == tcfg == cache.DEFAULT_TRANSFORM_CFG


def test_cache_resume_skips_finished_patches(tmp_path, oracle_lib):
    from b4d import cache, synth

    patches = np.stack([synth.vol(8, 8, 8, seed=s) for s in range(4)])
    d = str(tmp_path / "val")
    calls = []
    base = _oracle_targets(oracle_lib)

    def flaky(raw_u16, offsets, sigma, max_count):
        calls.append(len(raw_u16))
        if len(calls) == 2:
            raise RuntimeError("interrupted")
        return base(raw_u16, offsets, sigma, max_count)

    fgf = _oracle_fg(oracle_lib)
    with pytest.raises(RuntimeError):
        cache.write_patch_cache(d, patches, 37.0, 24.0, targets_fn=flaky, fg_fn=fgf, batch=2, split="val")
    first = np.load(os.path.join(d, "teacher.npy"))[:2].copy()
    assert not os.path.exists(os.path.join(d, "transform.json"))  # stamped last: cache not loadable yet
    with pytest.raises(ValueError):
        cache.load_patch_cache(d)
    n = cache.write_patch_cache(d, patches, 37.0, 24.0, targets_fn=base, fg_fn=fgf, batch=2, split="val")
    assert n == 2  # only the unfinished half
    raw, teacher, fgm, _ = cache.load_patch_cache(d)
    assert np.array_equal(teacher[:2], first) and np.abs(teacher[2:]).sum() > 0
    # no annotation mask supplied: the reference's fallback, make_foreground_mask(raw) (data_handling.py:444)
    assert np.array_equal(fgm.astype(bool), fgf(patches, np.full(4, 37.0, np.float32)))
    # a different configuration starts over
    assert cache.write_patch_cache(d, patches, 37.0, 10.0, targets_fn=base, fg_fn=fgf, batch=4) == 4
    # ... and so do different offsets, a different clip or different patches with the same N, shape and sigma
    # (the resume fingerprint in config.json); the same input again computes nothing
    assert cache.write_patch_cache(d, patches, 37.0, 10.0, targets_fn=base, fg_fn=fgf, batch=4) == 0
    assert cache.write_patch_cache(d, patches, 36.0, 10.0, targets_fn=base, fg_fn=fgf, batch=4) == 4
    assert cache.write_patch_cache(d, patches, 36.0, 10.0, targets_fn=base, fg_fn=fgf, batch=4, max_count=1000.0) == 4
    other = patches.copy()
    other[2, 3, 3, 3] ^= 1
    assert cache.write_patch_cache(d, other, 36.0, 10.0, targets_fn=base, fg_fn=fgf, batch=4, max_count=1000.0) == 4


def test_cache_rejects_bad_input(tmp_path):
    from b4d import cache

    with pytest.raises(ValueError):
        cache.write_patch_cache(str(tmp_path / "x"), np.zeros((2, 8, 8, 8), np.float32), 0.0, 24.0)
    with pytest.raises(ValueError):
        cache.write_patch_cache(str(tmp_path / "y"), np.zeros((2, 8, 8, 8), np.uint16), 0.0, 24.0,
                                fg=np.zeros((2, 8, 8, 4), np.uint8))
