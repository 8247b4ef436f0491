"""Test infrastructure (never imported by the product).  Helpers shared by the parity tests and by the cpu_baseline
leg of bench.py: tracing float32-vs-float64 differences to flipped stage-2 matches."""
import numpy as np

from oracle import np_oracle


def flipped_support(shape, lists_a, lists_b, Ns=11):
    """Boolean mask of the voxels covered by any grouped block of a reference block whose stage-2 match
    list differs between two pipelines (lists = (widx[R, K], cnt[R]) from Oracle.stage2_matches), and the
    number of such reference blocks.  Block matching is a discontinuous decision: where the float32 and
    the float64 basic estimates round to different matching images a near-tied candidate can enter or
    leave a group; the output can only differ by more than rounding noise under such a group."""
    (wa, ca), (wb, cb) = lists_a, lists_b
    K = wa.shape[1]
    valid_a = np.arange(K)[None, :] < ca[:, None]
    valid_b = np.arange(K)[None, :] < cb[:, None]
    flipped = (ca != cb) | ((wa != wb) & valid_a & valid_b).any(1)
    mask = np.zeros(shape, dtype=bool)
    rz, ry, rx = (np_oracle.ref_origins(n) for n in shape)
    r = Ns // 2
    for ri in np.flatnonzero(flipped):
        oz = rz[ri // (len(ry) * len(rx))]
        oy = ry[(ri // len(rx)) % len(ry)]
        ox = rx[ri % len(rx)]
        for w, valid in ((wa[ri], valid_a[ri]), (wb[ri], valid_b[ri])):
            for wi in w[valid]:
                wi = int(wi)
                cz, cy, cx = oz - r + wi // (Ns * Ns), oy - r + (wi // Ns) % Ns, ox - r + wi % Ns
                mask[cz : cz + 4, cy : cy + 4, cx : cx + 4] = True
    return mask, int(flipped.sum())


def check_against_f64(y, f64_out, lists_y, lists_f64, max_abs=0.5, quiet_abs=0.05, Ns=11):
    """The north-star bar, stated exactly: |y - f64| <= max_abs everywhere EXCEPT under a flipped stage-2
    match (proved voxel by voxel), and <= quiet_abs away from any flip.  Returns a report dict."""
    d = np.abs(y.astype(np.float64) - f64_out.astype(np.float64))
    mask, nflip = flipped_support(y.shape, lists_y, lists_f64, Ns)
    big = d > max_abs
    outside = float(d[~mask].max()) if (~mask).any() else 0.0
    report = {"max_abs": float(d.max()), "max_abs_outside_flips": outside, "flipped_groups": nflip,
              "voxels_over_bar": int(big.sum()), "voxels_over_bar_not_under_a_flip": int((big & ~mask).sum()),
              "voxels_under_flips": int(mask.sum())}
    assert report["voxels_over_bar_not_under_a_flip"] == 0, report
    assert outside <= quiet_abs, report
    return report
