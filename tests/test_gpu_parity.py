"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on B200.

Bars (BASELINE.json north_star / DESIGN.md §5):
  * stage-1 match lists: bit-exact (idx, ssd, count);
  * denoised output: BIT-EXACT vs the oracle's float32 mirror (same discrete
    decisions, same rounding; aggregation is always order-independent fixed point,
    the `deterministic` profile flag is kept for compatibility);
  * denoised output vs the plain float64 oracle: rel-L2 <= 1e-3 and max-abs <= 0.5 (north_star).
    Block matching is discontinuous, so a float64 pipeline may flip a handful of near-tied
    stage-2 matches; the tests extract the stage-2 match lists of both pipelines and PROVE that
    every voxel above 0.5 lies under a flipped group, and that the difference stays below 0.01
    everywhere else (oracle/parity_util.py);
  * quantize, statistics: bit-exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAX_ABS = 0.5
REL_L2 = 1e-3


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def b4d_mod():
    import b4d

    return b4d


@pytest.fixture(scope="module")
def dn(b4d_mod):
    d = b4d_mod.Denoiser(0)
    yield d
    d.close()


def _volumes():
    from b4d import synth

    rng = np.random.default_rng(0)
    return {
        "synth_ragged": synth.vol(21, 26, 31, seed=3),
        "synth_40": synth.vol(40, 40, 40, seed=5),
        "noise": np.clip(rng.normal(100, 24, (16, 17, 18)), 0, 65535).astype(np.uint16),
        "constant": np.full((12, 12, 12), 77, np.uint16),
        "single_block": np.clip(rng.normal(100, 24, (4, 4, 4)), 0, 65535).astype(np.uint16),
        "thin": synth.vol(4, 9, 33, seed=8),
        "bright_wide_range": (rng.integers(0, 2, (20, 20, 20)) * 60000 + rng.integers(0, 50, (20, 20, 20))).astype(
            np.uint16
        ),
        "ramp": (np.arange(18 * 18 * 18).reshape(18, 18, 18) % 4096).astype(np.uint16),
    }


@pytest.mark.parametrize("name", sorted(_volumes()))
def test_stage1_match_lists_bit_exact(name, dn, oracle_lib):
    vol = _volumes()[name]
    for sigma in (24.0, 10.0):
        gi, gs, gc = dn.match_stage1(vol, sigma)
        oi, os_, oc = oracle_lib.Oracle("f64").match_stage1(vol, sigma)
        assert np.array_equal(gc, oc)
        assert np.array_equal(gi, oi)
        assert np.array_equal(gs, os_)


def test_stage1_match_lists_config3_slice(dn, oracle_lib):
    """BASELINE config 3 is a 256^3 tile; the oracle finishes a 96^3 sub-tile of the
    same seeded volume in seconds — same code path, R = 32 768 reference blocks."""
    from b4d import synth

    vol = synth.vol(96, 96, 96, seed=3)
    gi, gs, gc = dn.match_stage1(vol, 24.0)
    oi, os_, oc = oracle_lib.Oracle("f64").match_stage1(vol, 24.0)
    assert gi.shape == (32 ** 3, 16)
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi) and np.array_equal(gs, os_)


def test_stage1_match_lists_config3_full_size(dn, oracle_lib):
    """BASELINE config 3 at its stated size: the 256^3 tile, R = 614 125 reference blocks, K = 16 —
    idx[R, K], ssd[R, K], count[R] bit-exact against the oracle's brute-force matcher (OpenMP: seconds)."""
    from b4d import synth

    vol = synth.vol(256, 256, 256, seed=3)
    gi, gs, gc = dn.match_stage1(vol, 24.0)
    oi, os_, oc = oracle_lib.Oracle("f64").match_stage1(vol, 24.0)
    assert gi.shape == (85 ** 3, 16)
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi) and np.array_equal(gs, os_)


@pytest.mark.parametrize("ns,k", [(7, 8), (15, 16), (5, 4), (13, 32)])
def test_stage1_other_windows(ns, k, b4d_mod, oracle_lib):
    from b4d import synth

    vol = synth.vol(24, 23, 22, seed=4)
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(search_window_ht=(ns // 2,) * 3, max_stack_size_ht=k))
    gi, gs, gc = d.match_stage1(vol, 24.0)
    oi, os_, oc = oracle_lib.Oracle("f64", search_ht=ns, k_ht=k).match_stage1(vol, 24.0)
    d.close()
    assert np.array_equal(gc, oc) and np.array_equal(gi, oi) and np.array_equal(gs, os_)


@pytest.mark.parametrize("name", ["synth_ragged", "synth_40", "noise", "constant", "single_block", "thin", "bright_wide_range"])
def test_deterministic_mode_bit_exact_vs_mirror(name, b4d_mod, oracle_lib):
    vol = _volumes()[name]
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(deterministic=True))
    for stages in (1, 2):
        d.set_profile(b4d_mod.BM4DProfile(deterministic=True), stages)
        y = d.denoise(vol, 24.0)
        m = oracle_lib.Oracle("mirror", stages=stages).denoise(vol, 24.0)
        assert y.dtype == np.float32 and y.shape == vol.shape
        assert np.array_equal(y, m), "stage %d: max-abs %g" % (stages, np.abs(y - m).max())
    d.close()


def test_deterministic_f32_offset_input_bit_exact(b4d_mod, oracle_lib):
    """The precompute path: raw = uint16 -> float32 - offset (data_handling.py:353-354)."""
    vol = _volumes()["synth_40"]
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(deterministic=True))
    for off in (37.0, 36.37):
        raw = vol.astype(np.float32) - np.float32(off)
        assert np.array_equal(d.denoise(raw, 24.0), oracle_lib.Oracle("mirror").denoise(raw, 24.0))
    # non-integral float data takes the quantised-matching path on both sides
    rng = np.random.default_rng(1)
    z = rng.normal(0.4, 0.05, (14, 15, 16)).astype(np.float32)
    assert np.array_equal(d.denoise(z, 0.05), oracle_lib.Oracle("mirror").denoise(z, 0.05))
    d.close()


def test_default_profile_config1_patch(dn, oracle_lib):
    """BASELINE config 1: one 64^3 patch through the precompute path, default profile."""
    from b4d import synth

    vol = synth.vol(64, 64, 64, seed=1)  # BASELINE config 1 input
    raw = vol.astype(np.float32) - np.float32(37.0)
    from oracle import parity_util

    for z in (vol, raw):
        y = dn.denoise(z, 24.0)
        om, of = oracle_lib.Oracle("mirror"), oracle_lib.Oracle("f64")
        m = om.denoise(z, 24.0)
        f = of.denoise(z, 24.0)
        assert np.array_equal(y, m)
        assert rel_l2(y, f) <= REL_L2
        # max-abs <= 0.5 except under a flipped stage-2 match (the device output equals the mirror bit for
        # bit, so the mirror's match lists are the device's), <= 0.01 away from any flip
        rep = parity_util.check_against_f64(y, f, om.stage2_matches(z.shape), of.stage2_matches(z.shape),
                                            max_abs=MAX_ABS, quiet_abs=0.01)
        print("device vs f64:", rep)
    # evaluator surface (evaluate.py:201-202): non-contiguous uint16 view, sigma 10
    view = vol[5:-5, 5:-5, 5:-5]
    import b4d

    yv = b4d.bm4d(view, 10)
    mv = oracle_lib.Oracle("mirror").denoise(np.ascontiguousarray(view), 10.0)
    assert yv.shape == (54, 54, 54) and np.array_equal(yv, mv)
    assert np.array_equal(np.maximum(yv, 0).astype(int), np.maximum(mv, 0).astype(int))


def test_batch_equals_per_patch_and_precompute_targets(b4d_mod, oracle_lib):
    from b4d import synth

    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(deterministic=True))
    batch = np.stack([synth.vol(20, 22, 24, seed=s) for s in (1, 2, 3)])
    yb = d.denoise(batch, 24.0)
    for i in range(3):
        assert np.array_equal(yb[i], d.denoise(batch[i], 24.0))
        assert np.array_equal(yb[i], oracle_lib.Oracle("mirror").denoise(batch[i], 24.0))
    d.close()
    # one launch for the whole batch, per-patch offsets (b4d_targets_u16): equal, bit for bit, to the float32
    # entry point on each offset-subtracted patch followed by the clip (data_handling.py:332-333, :353-354)
    offs = [37.0, 36.37, 12.5]
    raw, teacher = b4d_mod.precompute_targets(batch, offs, 24.0)
    assert raw.dtype == np.float32 and teacher.dtype == np.float32  # scripts/precompute.py:204-213
    assert teacher.min() >= 0.0 and teacher.max() <= 65535.0
    o = oracle_lib.Oracle("mirror")
    for i in range(3):
        ri = oracle_lib.read_counts(batch[i], np.float32(offs[i]))
        assert np.array_equal(raw[i], ri)
        assert np.array_equal(teacher[i], np.clip(o.denoise(ri, 24.0), 0, 65535))
        assert np.array_equal(teacher[i], np.clip(b4d_mod.bm4d(ri, 24.0), 0, 65535))
    # four chunks on two buffer sets (host transfers pipelined against the kernels): same bytes
    big = np.concatenate([batch] * 4)[:11]
    boffs = np.resize(np.asarray(offs, np.float32), 11)
    h = b4d_mod.get_denoiser()
    h.set_pipeline_min_voxels(0)
    praw, pteach = b4d_mod.precompute_targets(big, boffs, 24.0)
    h.set_pipeline_min_voxels(1 << 26)
    for i in range(11):
        assert np.array_equal(praw[i], raw[i % 3]) and np.array_equal(pteach[i], teacher[i % 3])
    # a patch size that is not a multiple of 8 voxels (the scalar form of the offset kernel)
    odd = np.stack([synth.vol(9, 10, 11, seed=s) for s in (5, 6, 7)])
    oraw, oteach = b4d_mod.precompute_targets(odd, [1.5, 37.0, 36.37], 24.0)
    for i, o in enumerate((1.5, 37.0, 36.37)):
        ri = oracle_lib.read_counts(odd[i], np.float32(o))
        assert np.array_equal(oraw[i], ri) and np.array_equal(oteach[i], np.clip(b4d_mod.bm4d(ri, 24.0), 0, 65535))
    _, low = b4d_mod.precompute_targets(batch[:1], 37.0, 24.0, max_count=150.0)
    assert low.max() <= 150.0 and np.array_equal(low[0], np.minimum(teacher[0], 150.0))


def test_slabs_equal_whole_on_device(b4d_mod):
    """Shard == whole, bit for bit, in deterministic mode."""
    from b4d import synth
    from b4d.sharding import halo_planes, slab_plan

    vol = synth.vol(96, 24, 28, seed=9)
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(deterministic=True))
    whole = d.denoise(vol, 24.0)
    halo = halo_planes(11, 11, 2)
    for world in (2, 3):
        parts = []
        for rank in range(world):
            ob, oe, zb, ze = slab_plan(96, world, rank, halo)
            parts.append(d.denoise_slab(vol[zb:ze], zb, 96, ob, oe, 24.0))
        assert np.array_equal(np.concatenate(parts, 0), whole)
    d.close()


def test_torch_device_tensors_in_out(dn, oracle_lib):
    import torch

    from b4d import synth

    vol = synth.vol(24, 24, 24, seed=2)
    t = torch.from_numpy(vol).cuda()
    y = dn.denoise(t, 24.0)
    assert y.is_cuda and y.dtype == torch.float32 and tuple(y.shape) == vol.shape
    assert np.abs(y.cpu().numpy() - oracle_lib.Oracle("mirror").denoise(vol, 24.0)).max() <= MAX_ABS
    q = dn.quantize(y)
    assert q.is_cuda and q.dtype == torch.uint16
    assert np.array_equal(q.cpu().numpy(), oracle_lib.quantize_reference(y.cpu().numpy()))


def test_quantize_bit_exact(dn, oracle_lib):
    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "reference_vectors.npz"))
    x = gold["quant_x"]
    for i, off in enumerate(gold["quant_offsets"]):
        assert np.array_equal(dn.quantize(x, 0.0, float(off), 1.0), gold["quant_q%d" % i])  # reference output
    rng = np.random.default_rng(2)
    x = rng.normal(500, 700, 1_000_003).astype(np.float32)
    assert np.array_equal(dn.quantize(x), oracle_lib.quantize_reference(x))
    for osub, oadd, step in ((3.5, 37.0, 2.5), (0.0, 0.0, 12.0), (36.37, 0.0, 1.0), (0.0, 37.0, 17.79)):
        assert np.array_equal(dn.quantize(x, osub, oadd, step), oracle_lib.quantize_noise_scaled(x, osub, oadd, step))
    assert dn.quantize(np.zeros(0, np.float32)).size == 0
    with pytest.raises(ValueError):
        dn.quantize(x, step=0.5)
    # the evaluator's truncating variant (evaluate.py:202 + the uint16 cast of img_util.py:420-423), written
    # with the reference's own NumPy expressions; values above 65535 wrap like the NumPy cast
    xt = np.concatenate([x, np.float32([0.0, -0.0, 0.999, 1.0, 65535.9, 65536.0, 70000.5, -3.2])])
    assert np.array_equal(dn.quantize(xt, truncate=True), oracle_lib.quantize_truncating(xt))
    assert np.array_equal(dn.quantize(xt, truncate=True),
                          np.ascontiguousarray(np.maximum(xt, 0).astype(int), dtype=np.uint16))
    assert np.array_equal(dn.quantize(xt, 3.5, 37.0, 2.5, truncate=True),
                          oracle_lib.quantize_truncating((xt - np.float32(3.5) + np.float32(37.0)) / np.float32(2.5)))


def test_fused_denoise_quantize_equals_the_two_calls(dn, b4d_mod, oracle_lib):
    """K6 + K7 fused (b4d_denoise_q16_u16 and the slab forms): the uint16 volume must equal K7 applied to the
    float32 output of the unfused call, bit for bit — rounding and truncating modes, offsets, noise-scaled
    step; device tensors and host arrays (pipelined copy-out); and the oracle's restatement."""
    import torch

    from b4d import synth

    vol = synth.vol(40, 44, 48, seed=11)
    y = dn.denoise(vol, 24.0)
    for osub, oadd, step, trunc in ((0.0, 0.0, 1.0, False), (36.5, 0.0, 1.0, False), (36.5, 3.0, 12.6, False),
                                    (0.0, 0.0, 1.0, True), (30.0, 0.0, 2.5, True)):
        want = dn.quantize(y, osub, oadd, step, truncate=trunc)
        got = dn.denoise_quantized(vol, 24.0, osub, oadd, step, truncate=trunc)
        assert got.dtype == np.uint16 and np.array_equal(got, want)
        gt = dn.denoise_quantized(torch.from_numpy(vol).cuda(), 24.0, osub, oadd, step, truncate=trunc)
        assert gt.dtype == torch.uint16 and np.array_equal(gt.cpu().numpy(), want)
    o = oracle_lib.Oracle("mirror")
    assert np.array_equal(dn.denoise_quantized(vol, 24.0, 36.5, 0.0, 2.0),
                          oracle_lib.quantize_noise_scaled(o.denoise(vol, 24.0), 36.5, 0.0, 2.0))
    # slab form: owned planes of a haloed slab, and the pipelined host path on a volume large enough for it
    big = synth.vol(160, 96, 96, seed=12)
    dn.set_pipeline_min_voxels(1 << 20)
    try:
        yb = dn.denoise(big, 24.0)
        qb = dn.denoise_quantized(big, 24.0, 36.5, 0.0, 1.0)
        assert np.array_equal(qb, dn.quantize(yb, 36.5, 0.0, 1.0))
        qs = dn.denoise_slab(big[20:130], 20, 160, 46, 104, 24.0, quantize=(36.5, 0.0, 1.0))
        assert np.array_equal(qs, qb[46:104])
    finally:
        dn.set_pipeline_min_voxels(1 << 26)


def test_tile_stats_exact(dn, oracle_lib):
    from b4d import synth
    from b4d.sharding import stats_from_hist

    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "reference_vectors.npz"))
    data, pos = gold["offset_data"], 0
    for n, p, want in zip(gold["offset_n"], gold["offset_pct"], gold["offset_val"]):
        s = data[pos : pos + n]
        pos += n
        assert dn.tile_stats(s, float(p))["offset"] == want  # reference estimate_offset output
    vol = synth.vol(64, 64, 64, seed=1)
    st, hist = dn.tile_stats(vol, 0.1, return_hist=True)
    assert np.array_equal(hist, np.bincount(vol.reshape(-1), minlength=65536))
    med, mad, sigma = oracle_lib.robust_sigma(vol)
    assert (st["median"], st["mad"], st["sigma"]) == (med, mad, sigma)
    assert st["offset"] == oracle_lib.estimate_offset(vol, 0.1)
    assert st == stats_from_hist(hist, 0.1)


def test_errors(dn, b4d_mod):
    with pytest.raises(ValueError):
        dn.denoise(np.zeros((3, 8, 8), np.uint16), 24.0)
    with pytest.raises(ValueError):
        dn.denoise(np.zeros((8, 8, 8), np.uint16), -1.0)
    with pytest.raises(NotImplementedError):
        dn.denoise(np.zeros((8, 8, 8), np.uint16), 500.0)  # tau*sigma^2*64 exceeds the 32-bit key
    with pytest.raises(ValueError):
        b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(max_stack_size_ht=12))
    # non-finite float input is refused (NaN as well as infinity: fmin / fmax reductions alone drop NaNs)
    for bad in (np.nan, np.inf, -np.inf):
        z = np.full((8, 9, 10), 100.0, np.float32)
        z[3, 4, 5] = bad
        with pytest.raises(ValueError):
            dn.denoise(z, 24.0)


def test_float_input_with_a_large_dc_level(dn, oracle_lib):
    """float32 data far from zero (a DC level much larger than the range) would push the fixed-point numerator
    terms past their clamp; such inputs are denoised on z - c0 and c0 is added back.  Bit-exact against the
    mirror (which states the same rule), within the float32 spacing of the DC level against the float64 oracle,
    and equal to the DC-free result up to that spacing (medians; a near-tied match may flip)."""
    from b4d import synth

    v = synth.vol(24, 26, 28, seed=2).astype(np.float32)
    base = dn.denoise(v, 24.0)
    for dc in (1.0e6, 3.0e7, -2.5e5):
        z = (v + np.float32(dc)).astype(np.float32)
        y = dn.denoise(z, 24.0)
        assert np.array_equal(y, oracle_lib.Oracle("mirror").denoise(z, 24.0))
        f = oracle_lib.Oracle("f64").denoise(z, 24.0)
        ulp = float(np.spacing(np.float32(abs(dc) + 65536.0)))
        assert np.abs(y.astype(np.float64) - f).max() <= max(ulp, 0.5)
        # against the DC-free result: the same image up to float32 spacing, except under a flipped near-tied match
        if ulp < 0.1:  # (at 3e7 the float32 spacing is 2: adding the DC level already changed the data)
            delta = np.abs((y.astype(np.float64) - dc) - base)
            assert np.median(delta) <= 2 * ulp + 1e-3 and delta.max() <= 4.0
    x = np.random.default_rng(0).normal(1000.0, 0.05, (20, 21, 22)).astype(np.float32)  # non-integral, sigma << DC
    y = dn.denoise(x, 0.05)
    assert np.array_equal(y, oracle_lib.Oracle("mirror").denoise(x, 0.05))
    assert np.std(y - 1000.0) < 0.5 * np.std(x - 1000.0)


def test_full_size_properties_128(dn, b4d_mod):
    """BASELINE config 2's patch size (128^3): properties that need no oracle —
    constant volume is a fixed point, output finite, noise reduced."""
    from b4d import synth

    vol = synth.vol(128, 128, 128, seed=1000)
    clean = synth.clean_vol(128, 128, 128, 1000)
    y = dn.denoise(vol, 24.0)
    assert np.isfinite(y).all()
    assert np.sqrt(np.mean((y - clean) ** 2)) < 0.5 * np.sqrt(np.mean((vol - clean) ** 2))
    # constant volume: every SSD ties at 0, every group takes the 32 lowest-index candidates, so
    # some voxels collect thousands of contributions — float atomics drift by ~1e-5 relative,
    # the fixed-point aggregation does not
    c = np.full((128, 128, 128), 4321, np.uint16)
    assert np.abs(dn.denoise(c, 24.0) - 4321.0).max() < 0.25
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(deterministic=True))
    assert np.abs(d.denoise(c, 24.0) - 4321.0).max() < 1e-2
    d.close()


def test_patch_cache_written_by_the_gpu_path(tmp_path, b4d_mod, oracle_lib):
    """the cache precompute.py writes (raw/teacher/fg .npy + transform.json),
    filled by batched GPU launches; teacher == clip(oracle bm4d(raw), 0, 65535) bit for bit."""
    from b4d import cache, synth

    patches = np.stack([synth.vol(20, 20, 20, seed=40 + s) for s in range(5)])
    offs = np.array([37.0, 37.0, 12.5, 37.0, 0.0], np.float32)
    d = str(tmp_path / "train")
    assert cache.write_patch_cache(d, patches, offs, 24.0, batch=3) == 5
    raw, teacher, fg, tcfg = cache.load_patch_cache(d)
    assert raw.dtype == np.float32 and teacher.dtype == np.float32 and fg.dtype == np.uint8
    o = oracle_lib.Oracle("mirror")
    for i in (0, 2, 4):
        want = np.clip(o.denoise(oracle_lib.read_counts(patches[i], float(offs[i])), 24.0), 0, 65535)
        assert np.array_equal(teacher[i], want)
        # no annotation mask supplied -> the reference's fallback make_foreground_mask(raw) (data_handling.py:444)
        assert np.array_equal(fg[i].astype(bool),
                              oracle_lib.make_foreground_mask_reference(oracle_lib.read_counts(patches[i], offs[i])))
    assert cache.write_patch_cache(d, patches, offs, 24.0, batch=3) == 0  # resume: nothing left


@pytest.mark.parametrize(
    "kw,okw",
    [
        # every filter-kernel instantiation: (stage, window class, group capacity)
        (dict(search_window_ht=(7, 7, 7), search_window_wiener=(7, 7, 7)), dict(search_ht=15, search_wie=15)),
        (dict(search_window_ht=(6, 6, 6), search_window_wiener=(3, 3, 3)), dict(search_ht=13, search_wie=7)),
        (dict(max_stack_size_ht=32, max_stack_size_wiener=16), dict(k_ht=32, k_wie=16)),
        (dict(max_stack_size_ht=8, max_stack_size_wiener=8), dict(k_ht=8, k_wie=8)),
        (dict(max_stack_size_ht=1, max_stack_size_wiener=2, search_window_ht=(1, 1, 1)), dict(k_ht=1, k_wie=2, search_ht=3)),
        (dict(beta=0.0, lambda_thr=3.1, tau_match_ht=1.5), dict(kaiser_beta=0.0, lambda_ht=3.1, tau_ht=1.5)),
    ],
)
def test_other_profiles_bit_exact_vs_mirror(kw, okw, b4d_mod, oracle_lib):
    """Non-default windows / group sizes / constants take other kernel instantiations (two-reference
    tiles for windows above 11, 32-slot hard-threshold groups, 16-slot Wiener groups without the
    warp-pair split); all must stay bit-exact against the mirror."""
    from b4d import synth

    rng = np.random.default_rng(5)
    vols = [synth.vol(22, 25, 28, seed=21), np.clip(rng.normal(300, 24, (17, 18, 20)), 0, 65535).astype(np.uint16)]
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(**kw))
    for vol in vols:
        for stages in (1, 2):
            d.set_profile(b4d_mod.BM4DProfile(**kw), stages)
            y = d.denoise(vol, 24.0)
            m = oracle_lib.Oracle("mirror", stages=stages, **okw).denoise(vol, 24.0)
            assert np.array_equal(y, m), "stages %d: max-abs %g" % (stages, np.abs(y - m).max())
    d.close()


def test_many_contributions_do_not_overflow_the_shared_accumulators(b4d_mod, oracle_lib):
    """A constant volume makes every SSD tie at 0: each group takes the lowest-index candidates, the
    grouped blocks overlap as much as they can and a voxel collects thousands of terms — the bound
    the 32-bit shared-memory limbs of the filter kernels are sized for."""
    c = np.full((40, 40, 40), 60000, np.uint16)
    d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(max_stack_size_ht=32))
    y = d.denoise(c, 24.0)
    m = oracle_lib.Oracle("mirror", k_ht=32).denoise(c, 24.0)
    assert np.array_equal(y, m) and np.abs(y - 60000.0).max() < 1e-2
    d.close()


def test_exchange_variant_slabs_equal_whole(b4d_mod):
    """SURVEY §8e "one exchange step": 13-plane halos, neighbours swap their exact basic-estimate
    planes between the stages.  Ranks are emulated by one handle each on this GPU; the union of
    the owned planes must equal the whole-volume result bit for bit."""
    from b4d import synth
    from b4d.sharding import exchange_halo, slab_plan

    vol = synth.vol(96, 24, 28, seed=9)
    whole = b4d_mod.Denoiser(0).denoise(vol, 24.0)
    h = exchange_halo(11, 11)
    for world in (2, 3):
        ranks = []
        for r in range(world):
            ob, oe, zb, ze = slab_plan(96, world, r, h)
            d = b4d_mod.Denoiser(0)
            d.slab_stage1(vol[zb:ze], zb, 96, 24.0)
            ranks.append((d, ob, oe, zb, ze))
        sends = []
        for d, ob, oe, zb, ze in ranks:  # what every rank would send: exact planes next to its faces
            sends.append((d.slab_basic(ob - zb, h), d.slab_basic(oe - zb - h, h)))
        parts = []
        for r, (d, ob, oe, zb, ze) in enumerate(ranks):
            if r > 0:
                d.slab_set_basic(0, sends[r - 1][1])        # the lower neighbour's top planes
            if r < world - 1:
                d.slab_set_basic(oe - zb, sends[r + 1][0])  # the upper neighbour's bottom planes
            parts.append(d.slab_stage2(ob, oe))
            d.close()
        assert np.array_equal(np.concatenate(parts, 0), whole)


def test_random_shapes_and_contents_equal_the_oracle(dn, oracle_lib):
    """Fuzz: 40 random shapes (4 .. 43 per axis, every fifth with rows that satisfy the TMA stride rule), noise on a
    random pedestal, bright blobs (byte, general and wide tiles side by side), flat halves (ties): stage-1 match
    lists and the two-stage output equal the oracle bit for bit.  (tools/fuzz_shapes.py runs the same at any size.)"""
    rng = np.random.default_rng(3)
    for c in range(40):
        shape = tuple(int(x) for x in rng.integers(4, 44, 3))
        if c % 5 == 0:
            shape = shape[:2] + (int(rng.choice([8, 16, 24, 32, 40, 48])),)
        kind = c % 4
        base = rng.integers(0, 3000)
        vol = base + rng.normal(0, 24.0, shape)
        if kind >= 1:
            z, y, x = (int(rng.integers(0, n)) for n in shape)
            vol[max(z - 3, 0) : z + 4, max(y - 3, 0) : y + 4, max(x - 3, 0) : x + 4] += float(rng.choice([400, 5000, 40000]))
        if kind == 3:
            vol[:, :, : shape[2] // 2] = base
        vol = np.clip(np.rint(vol), 0, 65535).astype(np.uint16)
        sigma = float(rng.choice([10.0, 24.0, 40.0]))
        o = oracle_lib.Oracle("mirror")
        gi, gs, gc = dn.match_stage1(vol, sigma)
        oi, os_, oc = o.match_stage1(vol, sigma)
        assert np.array_equal(gc, oc), (c, shape)
        valid = np.arange(gi.shape[1])[None, :] < gc[:, None]
        assert np.array_equal(gi[valid], oi[valid]) and np.array_equal(gs[valid], os_[valid]), (c, shape)
        assert np.array_equal(dn.denoise(vol, sigma), o.denoise(vol, sigma)), (c, shape)


def test_tma_and_fallback_staging_give_the_same_bytes(dn, b4d_mod):
    """Every TMA-fed kernel (K0, byte-matcher window, general-matcher table, normalise) has a cp.async / plain-load
    twin for volumes whose row pitch breaks the 16-byte rule; B4D_NO_TMA forces the twins on an aligned volume."""
    import os

    from b4d import synth

    vol = synth.vol(48, 40, 64, seed=11)
    vol[10:30, 8:30, 16:50] += 3000  # a bright structure: wide and general tiles next to byte tiles
    y = dn.denoise(vol, 24.0)
    q = dn.denoise_quantized(vol, 24.0, offset_sub=30.0, step=2.0)
    idx, ssd, cnt = dn.match_stage1(vol, 24.0)
    os.environ["B4D_NO_TMA"] = "1"
    try:
        y2 = dn.denoise(vol, 24.0)
        q2 = dn.denoise_quantized(vol, 24.0, offset_sub=30.0, step=2.0)
        idx2, ssd2, cnt2 = dn.match_stage1(vol, 24.0)
    finally:
        del os.environ["B4D_NO_TMA"]
    assert np.array_equal(y, y2) and np.array_equal(q, q2)
    assert np.array_equal(idx, idx2) and np.array_equal(ssd, ssd2) and np.array_equal(cnt, cnt2)


def test_overlapped_exchange_form_equals_whole(b4d_mod):
    """The device form of the exchange variant: after stage 1 every rank launches the part of the stage-2 front end
    that needs owned planes only (b4d_slab_stage2_begin, no wait) while the neighbours' planes are written straight
    into its basic-estimate buffer (b4d_slab_basic_ptr, zero-copy torch views; here a device-to-device copy stands
    in for the NCCL point-to-point).  Union of the owned planes == whole volume, bit for bit — float32 and the
    fused uint16 quantizer."""
    import torch

    from b4d import synth
    from b4d.sharding import exchange_halo, slab_plan

    dev = torch.device("cuda", 0)
    vol = synth.vol(120, 40, 36, seed=19)
    whole_dn = b4d_mod.Denoiser(0)
    whole = whole_dn.denoise(vol, 24.0)
    whole_q = whole_dn.quantize(whole, 36.5, 0.0, 2.0)
    h = exchange_halo(11, 11)
    for world in (2, 3):
        ranks = []
        for r in range(world):
            ob, oe, zb, ze = slab_plan(120, world, r, h)
            d = b4d_mod.Denoiser(0)
            d.slab_stage1(torch.from_numpy(vol[zb:ze]).to(dev), zb, 120, 24.0)
            ranks.append((d, ob, oe, zb, ze, d.slab_basic_tensor(dev)))
        torch.cuda.synchronize()
        sends = [(b[ob - zb : ob - zb + h].clone(), b[oe - zb - h : oe - zb].clone()) for d, ob, oe, zb, ze, b in ranks]
        for d, ob, oe, zb, ze, b in ranks:
            d.slab_stage2_begin(ob, oe)  # runs while the planes below are "exchanged"
        for r, (d, ob, oe, zb, ze, b) in enumerate(ranks):
            if r > 0:
                b[0 : ob - zb].copy_(sends[r - 1][1])
            if r < world - 1:
                b[oe - zb : oe - zb + h].copy_(sends[r + 1][0])
        torch.cuda.synchronize()
        parts, qparts = [], []
        for r, (d, ob, oe, zb, ze, b) in enumerate(ranks):
            if r % 2 == 0:
                parts.append(d.slab_stage2(ob, oe, device=dev).cpu().numpy())
                qparts.append(None)
            else:
                qparts.append(d.slab_stage2(ob, oe, device=dev, quantize=(36.5, 0.0, 2.0)).cpu().numpy())
                parts.append(None)
            d.close()
        for r, (d, ob, oe, zb, ze, b) in enumerate(ranks):
            if parts[r] is not None:
                assert np.array_equal(parts[r], whole[ob:oe]), "world %d rank %d" % (world, r)
            else:
                assert np.array_equal(qparts[r], whole_q[ob:oe]), "world %d rank %d (quantized)" % (world, r)


def test_pipelined_host_transfers_equal_device_result(b4d_mod):
    """With host buffers and a deep volume the transfers are pipelined: the uint16 input is uploaded
    in z chunks behind the stage-1 matcher, and stage 2 runs in z chunks whose finished planes are
    normalised and copied out on a second stream.  The result must equal the device-resident path
    (one upload, one launch per kernel) bit for bit — volume, slab and two-call form."""
    import torch

    from b4d import synth

    vol = synth.vol(208, 16, 20, seed=12)  # 69 reference planes: enough for 8 chunks
    d = b4d_mod.Denoiser(0)
    d.set_pipeline_min_voxels(0)  # pipelined at any size (default: only from 2^26 voxels up)
    dev = d.denoise(torch.from_numpy(vol).cuda(), 24.0).cpu().numpy()  # device in/out: unchunked
    assert np.array_equal(d.denoise(vol, 24.0), dev)                   # host in/out: chunked
    part = d.denoise_slab(vol[30:208], 30, 208, 60, 200, 24.0)         # slab, host out
    ref = d.denoise_slab(torch.from_numpy(vol[30:208]).cuda(), 30, 208, 60, 200, 24.0).cpu().numpy()
    assert np.array_equal(part, ref)
    d.slab_stage1(vol, 0, 208, 24.0)                                   # two-call form, host out
    assert np.array_equal(d.slab_stage2(0, 208), dev)
    d.close()


def test_pageable_and_pinned_host_arrays_give_the_same_bytes(dn, b4d_mod, oracle_lib):
    """NumPy arrays are pageable memory: they cross PCIe through the pinned ring + copy threads of
    the HostMover (16 MiB pieces, several per call, ragged tail); pinned torch tensors are copied
    directly.  Both must deliver every byte — checked on the bandwidth-bound quantize call with
    a buffer of many pieces, and on the denoiser with pageable against pinned buffers."""
    import torch

    from b4d import synth

    rng = np.random.default_rng(9)
    x = rng.normal(500, 700, 5 * (4 << 20) + 12_345).astype(np.float32)  # 5 pieces + a ragged one
    want = oracle_lib.quantize_reference(x)
    assert np.array_equal(dn.quantize(x), want)                                    # pageable in, pageable out
    xp = torch.from_numpy(x).pin_memory()
    assert np.array_equal(dn.quantize(xp).numpy(), want)                           # pinned in
    assert np.array_equal(dn.quantize(x[1:]), want[1:])                            # odd alignment
    st = dn.tile_stats(want)
    assert st["n"] == want.size and st["vmax"] == float(want.max())
    vol = synth.vol(208, 40, 52, seed=5)
    dn.set_pipeline_min_voxels(0)
    a = dn.denoise(vol, 24.0)                                                      # pageable, chunked pipeline
    b = dn.denoise(torch.from_numpy(vol).pin_memory(), 24.0).numpy()               # pinned, chunked pipeline
    dn.set_pipeline_min_voxels(1 << 26)
    assert np.array_equal(a, b) and np.array_equal(a, dn.denoise(vol, 24.0))       # == the unchunked path


@pytest.mark.parametrize(
    "shape,chunk",
    [((64, 64, 64), (64, 64, 64)), ((128, 64, 192), (64, 64, 64)), ((70, 66, 130), (64, 64, 64)),
     ((20, 33, 35), (8, 16, 4)), ((40, 36, 44), (32, 16, 8)), ((5, 3, 2), (64, 64, 64))],
)
def test_chunk_shuffle_bit_exact(shape, chunk, dn, oracle_lib):
    """K9 (first step): chunk gather + Blosc 2-byte shuffle + per-piece byte counts,
    against the restatement of compute_cratio's loop (utils/img_util.py:427-438) — bytes and counts equal."""
    import torch

    from b4d import synth

    vol = synth.vol(*shape, seed=sum(shape))
    vol[: shape[0] // 2] //= 8  # a region whose high bytes are all zero (the warp-uniform path)
    by, hist = oracle_lib.chunk_shuffle_reference(vol, chunk)
    gby, ghist = dn.chunk_shuffle(vol, chunk)
    assert gby.dtype == np.uint8 and np.array_equal(gby, by)
    assert ghist.dtype == np.uint32 and np.array_equal(ghist, hist)
    tby, thist = dn.chunk_shuffle(torch.from_numpy(vol).cuda(), chunk)  # device in -> device out
    assert tby.is_cuda and np.array_equal(tby.cpu().numpy(), by)
    assert np.array_equal(thist.cpu().numpy().view(np.uint32), hist)
    only_hist = dn.chunk_shuffle(vol, chunk, want_bytes=False)
    assert only_hist[0] is None and np.array_equal(only_hist[1], hist)


def test_cratio_of_shuffled_pieces_round_trips(dn, b4d_mod):
    """Device shuffle -> host zstd (the system library) -> decompress -> unshuffle gives the volume back;
    the ratio has the reference's shape (total / total, 2 decimals) and denoising raises it."""
    from b4d import codec, synth

    if not codec.zstd_available():
        pytest.skip("libzstd not loadable on this host")
    vol = synth.vol(70, 64, 128, seed=8)
    by, hist = dn.chunk_shuffle(vol)
    back = np.zeros_like(vol)
    for (z0, y0, x0), d, pos in codec.chunk_grid(vol.shape):
        nb = 2 * d[0] * d[1] * d[2]
        c = codec.zstd_compress(by[pos : pos + nb], 6)
        back[z0 : z0 + d[0], y0 : y0 + d[1], x0 : x0 + d[2]] = codec.unshuffle_piece(codec.zstd_decompress(c, nb), d)
    assert np.array_equal(back, vol)
    r_noisy = b4d_mod.compute_cratio(vol)
    den = b4d_mod.quantize(b4d_mod.bm4d(vol, 24.0))
    r_den = b4d_mod.compute_cratio(den)
    assert r_noisy == round(r_noisy, 2) and r_den > r_noisy > 1.0
    assert b4d_mod.estimate_cratio(dn.chunk_shuffle(den, want_bytes=False)[1]) > b4d_mod.estimate_cratio(hist)


def test_foreground_mask_matches_reference_golden(dn, b4d_mod, oracle_lib):
    """make_foreground_mask on the device (exact histogram medians, float32 statistic on the host,
    threshold + L1-ball dilation kernel) against the masks the live reference produced
    (tests/golden/reference_masks.npz), and a mixed batch against the restatement."""
    import os

    from b4d import synth

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_masks.npz"))
    for i in range(int(g["count"][0])):
        ci, off, k, dil = g["par%d" % i]
        u = g["u%d" % int(ci)]
        want = np.unpackbits(g["m%d" % i])[: u.size].astype(bool).reshape(u.shape)
        got = dn.foreground_mask(u, float(off), float(k), int(dil))
        assert got.dtype == np.bool_ and np.array_equal(got, want), (i, off, k, dil)
    batch = np.stack([synth.vol(15, 17, 19, seed=s) for s in range(40)])  # odd voxel count: unaligned patches
    batch[3, 4:9, 5:9, 6:12] += 3000
    offs = np.linspace(0.0, 40.0, 40).astype(np.float32)
    got = b4d_mod.make_foreground_mask(batch, offs)
    for i in (0, 3, 17, 39):
        assert np.array_equal(got[i], oracle_lib.make_foreground_mask_reference(oracle_lib.read_counts(batch[i], offs[i])))
    assert got[3].any()


def test_volume_larger_than_a_pass_is_slabbed_on_one_gpu(b4d_mod):
    """A uint16 volume that exceeds the per-pass scratch budget is denoised as consecutive z-slabs with
    halos on the same GPU (b4d_set_pass_voxels): same bytes as the one-pass result, timings summed."""
    from b4d import synth

    vol = synth.vol(100, 24, 28, seed=21)
    d = b4d_mod.Denoiser(0)
    whole = d.denoise(vol, 24.0)
    one = sum(v[1] for v in d.last_timings().values())
    d.set_pass_voxels(60 * 24 * 28)  # 60 planes per pass, 52 of them halo: 13 slabs of 8 planes
    assert np.array_equal(d.denoise(vol, 24.0), whole)
    assert sum(v[1] for v in d.last_timings().values()) > 5 * one
    d.set_pass_voxels(80 * 24 * 28)
    assert np.array_equal(d.denoise(vol, 24.0), whole)
    two = d.denoise(np.stack([vol, vol[::-1].copy()]), 24.0)  # a batch of oversize volumes: one after the other
    assert np.array_equal(two[0], whole) and np.array_equal(two[1], b4d_mod.bm4d(vol[::-1].copy(), 24.0))
    d.set_pass_voxels(40 * 24 * 28)  # two halos do not fit
    with pytest.raises(ValueError):
        d.denoise(vol, 24.0)
    d.set_pass_voxels(0)
    assert np.array_equal(d.denoise(vol, 24.0), whole)
    d.close()


def test_full_size_1024_translation_property(dn):
    """BASELINE config 4's size (1024^3, 2 GiB of uint16) through a property that needs no oracle: the
    volume is the seeded 128^3 tile repeated with period 128; the reference grid has step 3, so a shift
    by lcm(128, 3) = 384 voxels along every axis keeps both the data and the grid phase — the denoised
    interior must repeat with that shift, bit for bit, and equal the same region of a 384^3 run (whose
    search windows see identical data)."""
    import torch

    from b4d import synth

    tile = torch.from_numpy(synth.vol(128, 128, 128, seed=1000)).cuda()
    vol = tile.repeat(8, 8, 8)
    y = dn.denoise(vol, 24.0)
    assert y.shape == (1024, 1024, 1024) and y.dtype == torch.float32
    a = y[128:256, 128:256, 128:256]
    assert torch.equal(a, y[512:640, 512:640, 512:640]) and torch.equal(a, y[512:640, 128:256, 896 - 384 : 1024 - 384])
    small = dn.denoise(vol[:384, :384, :384].contiguous(), 24.0)
    assert torch.equal(a, small[128:256, 128:256, 128:256])
    assert bool(torch.isfinite(y).all())
    clean = torch.from_numpy(synth.clean_vol(128, 128, 128, 1000)).cuda()
    err_in = (tile.float() - clean).pow(2).mean().sqrt()
    err_out = (y[384:512, 384:512, 384:512] - clean).pow(2).mean().sqrt()
    assert float(err_out) < 0.5 * float(err_in)
    del y, vol, small
    torch.cuda.empty_cache()


def test_bench_volume_1024_crops_equal_the_oracle(dn, oracle_lib):
    """The ACTUAL benchmark input (bench.make_slab_device, seed 4, 1024^3) denoised whole on the device,
    compared with the oracle on haloed crops cut from it.  A crop whose origin is a multiple of the grid
    step (3) and whose side s has (s - 4) % 3 == 0 carries exactly the global reference grid; its result is
    exact at least 2 * (Ns - 1 + L - 1) = 26 voxels away from every cut face, so the
    interior must equal the whole-volume result BIT FOR BIT — in the middle of the volume, at the origin
    corner and at the far corner (where the crop face is the volume face and needs no halo)."""
    import torch

    import bench

    S, side, halo = 1024, 97, 27
    vol = bench.make_slab_device(S, 0, S, torch.device("cuda", 0))
    y = dn.denoise(vol, bench.SIGMA)
    assert y.shape == (S, S, S)
    o = oracle_lib.Oracle("mirror")
    for org in ((501, 300, 699), (0, 0, 0), (S - side, S - side, S - side), (0, 927, 402)):
        assert all(c % 3 == 0 for c in org) and (side - 4) % 3 == 0
        sl = tuple(slice(c, c + side) for c in org)
        m = o.denoise(vol[sl].cpu().numpy(), bench.SIGMA)
        inner = tuple(slice(0 if c == 0 else halo, side if c + side == S else side - halo) for c in org)
        got = y[sl][inner].cpu().numpy()
        assert got.size >= 43 ** 3
        assert np.array_equal(got, m[inner]), "crop at %r: max-abs %g" % (org, np.abs(got - m[inner]).max())
    del y, vol
    torch.cuda.empty_cache()


def test_coherence_gate_matches_the_reference(dn, b4d_mod, oracle_lib):
    """The sampler's spatial-coherence gate on the device (b4d_coherence_gate) against outputs of the LIVE reference
    (tests/golden/reference_coherence.npz): verdicts equal; per-segment voxel counts equal, lag autocorrelation and
    high-frequency energy fraction within 1e-9 (float64 sums in another order); a batch equals the single calls."""
    from test_oracle_golden import _coherence_cases

    cases = list(_coherence_cases())
    for lab, raw, kw, ids, scores, verdict in cases:
        rej, tabs = dn.coherence_scores(lab, raw, kw["smooth_sigma"], kw["coherence_lag"], kw["min_autocorr"],
                                        kw["max_highfreq_frac"], kw["min_segment_voxels"])
        assert bool(rej[0]) == verdict
        assert sorted(tabs[0]) == [int(v) for v in ids]
        for lid, want in zip(ids, scores):
            got = tabs[0][int(lid)]
            assert got[0] == int(want[0])
            assert abs(got[1] - want[1]) <= 1e-9 and abs(got[2] - want[2]) <= 1e-9, (lid, got, want)
        assert b4d_mod.patch_has_incoherent_segment(lab, raw, **kw) == verdict
    same = [c for c in cases if c[0].shape == (48, 48, 48) and c[2] == cases[0][2]]
    rej = dn.patch_has_incoherent_segment(np.stack([c[0] for c in same]), np.stack([c[1] for c in same]))
    assert rej.tolist() == [c[5] for c in same]
    # offsets do not matter (every statistic is shift invariant): raw counts minus a pedestal give the same verdict
    lab, raw, kw, ids, scores, verdict = cases[2]
    assert dn.patch_has_incoherent_segment(lab, raw + np.float32(1000.25), **kw) == verdict


def test_coloured_noise_psd_input(dn, b4d_mod, oracle_lib):
    """bm4d(z, sigma_psd) with an array PSD (the other half of the call surface): bit-exact against the mirror with
    the same per-coefficient variance tables, close to the float64 path, better than the white model on correlated
    noise; nu = 1 tables and a constant PSD reproduce the scalar path bit for bit; uint16 and float32 inputs, both
    group-size instantiations."""
    from test_oracle_bm4d import _coloured_case

    shape = (24, 28, 32)
    psd, clean, z = _coloured_case(shape)
    s, a, b = b4d_mod.noise_model_from_psd(psd)
    y = b4d_mod.bm4d(z, psd)
    om, of = oracle_lib.Oracle("mirror"), oracle_lib.Oracle("f64")
    om.set_noise_model(a, b)
    of.set_noise_model(a, b)
    assert np.array_equal(y, om.denoise(z, s))
    f = of.denoise(z, s)
    assert rel_l2(y, f) <= REL_L2 and np.abs(y - f).max() <= MAX_ABS
    white = b4d_mod.bm4d(z, s)
    rmse = lambda v: float(np.sqrt(np.mean((v - clean) ** 2)))  # noqa: E731
    assert rmse(y) < 0.85 * rmse(white)
    # the model is switched off again after the call; a constant PSD is the scalar path
    assert np.array_equal(b4d_mod.bm4d(z, s), white)
    assert np.array_equal(b4d_mod.bm4d(z, np.full(shape, s * s * np.prod(shape))), white)
    # nu = 1 tables == white, bit for bit; uint16 input; 16-block Wiener groups
    zu = np.clip(np.rint(z + 200.0), 0, 65535).astype(np.uint16)
    for kw, okw in (({}, {}), ({"max_stack_size_wiener": 16, "max_stack_size_ht": 8}, {"k_wie": 16, "k_ht": 8})):
        d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(**kw))
        w0 = d.denoise(zu, s)
        d.set_noise_model(np.ones(64, np.float32), np.ones(64, np.float32))
        assert np.array_equal(d.denoise(zu, s), w0)
        d.set_noise_model(a, b)
        o = oracle_lib.Oracle("mirror", **okw)
        o.set_noise_model(a, b)
        assert np.array_equal(d.denoise(zu, s), o.denoise(zu, s))
        d.close()
    with pytest.raises(NotImplementedError):
        d = b4d_mod.Denoiser(0, b4d_mod.BM4DProfile(search_window_ht=(7, 7, 7)))
        d.set_noise_model(a, b)
        d.denoise(zu, s)
