"""Host logic of the z-slab path: slab plan, histogram statistics and the one
collective (all-gather of per-slab histograms) under gloo, world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_plan_partitions_and_halos():
    from b4d.sharding import halo_planes, slab_plan

    assert halo_planes(11, 11, 2) == 26 and halo_planes(15, 15, 2) == 34 and halo_planes(11, 11, 1) == 13
    for z_total, world in ((1024, 8), (1000, 7), (64, 2), (30, 1)):
        covered = []
        for r in range(world):
            ob, oe, zb, ze = slab_plan(z_total, world, r, 26)
            assert 0 <= zb <= ob < oe <= ze <= z_total
            assert zb == max(0, ob - 26) and ze == min(z_total, oe + 26)
            covered.extend(range(ob, oe))
        assert covered == list(range(z_total))
    ob, oe, zb, ze = slab_plan(1024, 8, 3, 26)
    assert (oe - ob, ze - zb) == (128, 180)  # 128 + 2*26 planes on an interior rank
    with pytest.raises(ValueError):
        slab_plan(4, 8, 7, 26)


def test_stats_from_hist_equals_numpy():
    from b4d.sharding import stats_from_hist

    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 1000, 65537):
        x = np.clip(rng.normal(40, 25, n), 0, 65535).astype(np.uint16)
        x[rng.random(n) < 0.2] = 0
        xf = x.astype(np.float32)
        st = stats_from_hist(np.bincount(x, minlength=65536), 0.1)
        nz = xf[xf > 0]
        assert st["offset"] == float(np.percentile(nz if nz.size else xf, 0.1))
        med = np.median(xf)
        mad = np.median(np.abs(xf - med)) + 1e-6
        assert (st["median"], st["mad"], st["sigma"]) == (float(med), float(mad), float(1.4826 * mad))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from b4d import synth
    from b4d.sharding import halo_planes, merge_histograms, slab_plan, stats_from_hist
    from oracle import np_oracle

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        z_total = 36
        kw = dict(search_ht=5, search_wie=5, k_ht=8, k_wie=8)
        halo = halo_planes(5, 5, 2)
        ob, oe, zb, ze = slab_plan(z_total, world, rank, halo)
        # every rank regenerates only its slab of the seeded volume
        slab = synth.vol(ze - zb, 12, 12, seed=11, z0=zb)
        o = np_oracle.Oracle("mirror", **kw)  # CPU stand-in for the device in this host-logic test
        own = o.denoise_slab(slab, zb, z_total, ob, oe, 24.0)
        hist = torch.from_numpy(np.bincount(slab[ob - zb : oe - zb].reshape(-1), minlength=65536).astype(np.int64))
        total = merge_histograms(hist)
        st = stats_from_hist(total.numpy(), 1.0)
        q.put((rank, ob, oe, own, st))
    finally:
        dist.destroy_process_group()


def test_two_rank_slabs_and_histogram_allgather(oracle_lib):
    import torch.multiprocessing as mp

    from b4d import synth
    from b4d.sharding import stats_from_hist

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    vol = synth.vol(36, 12, 12, seed=11)
    whole = oracle_lib.Oracle("mirror", search_ht=5, search_wie=5, k_ht=8, k_wie=8).denoise(vol, 24.0)
    got = np.concatenate([r[3] for r in res], 0)
    assert np.array_equal(got, whole)  # shard == whole, bit for bit
    want = stats_from_hist(np.bincount(vol.reshape(-1), minlength=65536), 1.0)
    assert res[0][4] == want and res[1][4] == want  # every rank derives the same global statistics


def _xworker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "aind-exaspim-image-compression_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from b4d.sharding import exchange_planes

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        h, hw = 3, (4, 5)
        down = torch.full((h,) + hw, float(10 * rank + 1)) if rank > 0 else None
        up = torch.full((h,) + hw, float(10 * rank + 2)) if rank < world - 1 else None
        below = torch.zeros((h,) + hw) if rank > 0 else None
        above = torch.zeros((h,) + hw) if rank < world - 1 else None
        exchange_planes(down, up, below, above, rank, world)
        q.put((rank, None if below is None else float(below.mean()), None if above is None else float(above.mean())))
    finally:
        dist.destroy_process_group()


def test_neighbour_plane_exchange_three_ranks():
    """The exchange variant's one data-path collective: every rank swaps halo planes with its two
    neighbours (gloo here, NCCL point-to-point on the GPUs)."""
    import torch.multiprocessing as mp

    from b4d.sharding import exchange_halo, halo_planes, slab_plan

    assert exchange_halo(11, 11) == 13 and exchange_halo(15, 11) == 17 and 2 * exchange_halo(11, 11) == halo_planes(11, 11, 2)
    ob, oe, zb, ze = slab_plan(1024, 8, 3, exchange_halo(11, 11))
    assert (oe - ob, ze - zb) == (128, 154)  # 128 + 2*13 planes instead of 128 + 2*26
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_xworker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # rank r receives rank r-1's `up` (10(r-1)+2) from below and rank r+1's `down` (10(r+1)+1) from above
    assert res[0] == (0, None, 11.0) and res[1] == (1, 2.0, 21.0) and res[2] == (2, 12.0, None)


def test_library_statistics_of_a_merged_histogram_equal_the_numpy_form():
    """b4d_stats_from_hist (host arithmetic of libb4d, what the ranks evaluate on the summed histogram inside the
    step) against the NumPy form, which the golden tests pin to the reference's estimate_offset / robust sigma."""
    from b4d.sharding import stats_from_hist, stats_from_hist_lib

    rng = np.random.default_rng(3)
    for n, mu, pct in ((200001, 40.0, 0.1), (1000, 300.0, 1.0), (77777, 5.0, 50.0), (4096, 0.2, 1.0)):
        h = np.zeros(65536, np.int64)
        np.add.at(h, np.clip(rng.normal(mu, 24, n), 0, 65535).astype(np.int64), 1)
        a, b = stats_from_hist(h, pct), stats_from_hist_lib(h, pct)
        for k in ("n", "n_nonzero", "offset", "median", "mad", "sigma"):
            assert a[k] == b[k], (k, a, b)
