// b4d_match.cu — K1 / K4: exact-integer block matching on a uint16 image.
//
// Contract: for every reference block
// (4x4x4, origins on the step-3 grid plus the flush origin N-4) compute the
// exact SSD to every candidate origin of the Ns^3 window clipped to the volume,
// accept SSD <= tau, order by (SSD, window index), keep the first
// K' = 2^floor(log2(min(K, accepted))).
//
// Mapping: one CTA per 4x4x4 tile of reference blocks; the (Ns+12)^3 voxel
// neighbourhood of the tile is staged once in shared memory (zero filled outside
// the volume) and serves all 64 reference blocks.  One warp per reference
// block.  A lane owns one (dz, dy) row of the window at a time and slides along
// dx with Ns accumulators in registers: each shared-memory row of Ns+3 values
// feeds 4*Ns (sub, mad) pairs, the reference block lives in 64 registers.
// SSDs are uint32 when the tile's value range allows (64*(range)^2 < 2^32,
// checked per tile while staging), else uint64.
//
// Selection is exact and deterministic: key = SSD << KB | window index (unique),
// rejected candidates get 0xFFFFFFFF.  A running per-lane minimum (second
// minimum for K = 32) gives, through one 32-lane bitonic sort per iteration, an
// upper bound B with at least K keys <= B; only keys <= B are appended to a small
// per-warp survivor list, which is rank-sorted at the end.  If the list
// overflows (adversarial key order) the warp falls back to K rounds of
// "smallest key greater than the previous one", recomputing the SSDs.
#include "b4d_common.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int CAP = 256;  // survivor list entries per warp

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) { return __reduce_min_sync(B4D_FULL, v); }

// K-th smallest (0-based index kth) of one value per lane: 32-lane bitonic sort.
__device__ __forceinline__ uint32_t kth_smallest32(uint32_t v, int kth, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(B4D_FULL, v, j);
            const bool up = (lane & k) == 0;      // ascending half (k == 32: all ascending)
            const bool lower = (lane & j) == 0;
            const uint32_t mn = min(v, o), mx = max(v, o);
            v = (lower == up) ? mn : mx;
        }
    }
    return __shfl_sync(B4D_FULL, v, kth);
}

// SSDs of one (dz, dy) row of candidates, all NS dx positions, uint32 path.
template <int NS>
__device__ __forceinline__ void ssd_row_u32(const uint16_t *__restrict__ base, const int (&ref)[B4D_LV],
                                            uint32_t (&acc)[NS]) {
    constexpr int E = NS + 12;
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = 0u;
#pragma unroll
    for (int z = 0; z < 4; ++z) {
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            int v[NS + 3];
#pragma unroll
            for (int i = 0; i < NS + 3; ++i) v[i] = base[(z * E + y) * E + i];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int r = ref[(z * 4 + y) * 4 + x];
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const int d = v[x + j] - r;
                    acc[j] += (uint32_t)(d * d);
                }
            }
        }
    }
}

// Same, 64-bit accumulation for tiles whose value range is too wide for uint32.
// Rare (bright structures above 8191 counts over background); kept compact.
template <int NS>
__device__ __noinline__ void ssd_row_u64(const uint16_t *__restrict__ base, const uint16_t *__restrict__ refp,
                                         unsigned long long *acc) {
    constexpr int E = NS + 12;
    for (int j = 0; j < NS; ++j) acc[j] = 0ull;
#pragma unroll 1
    for (int z = 0; z < 4; ++z) {
#pragma unroll 1
        for (int y = 0; y < 4; ++y) {
            const uint16_t *row = base + (z * E + y) * E;
            const uint16_t *rr = refp + (z * E + y) * E;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int r = rr[x];
#pragma unroll 1
                for (int j = 0; j < NS; ++j) {
                    const int d = (int)row[x + j] - r;
                    const uint32_t ad = (uint32_t)(d < 0 ? -d : d);
                    acc[j] += (unsigned long long)ad * ad;
                }
            }
        }
    }
}

template <int NS, bool K32>
__global__ void __launch_bounds__(WARPS * 32, 2) k_match(const MatchParams p) {
    constexpr int R_ = NS / 2;
    constexpr int E = NS + 12;
    constexpr int KB = (NS * NS * NS <= 2048) ? 11 : 12;
    constexpr int UNITS = NS * NS;
    constexpr int ITERS = (UNITS + 31) / 32;
    constexpr int EV = E * E * E;

    extern __shared__ __align__(16) uint16_t s_win[];  // E^3 (+ pad)
    __shared__ uint32_t s_surv[WARPS][CAP];
    __shared__ int s_cnt[WARPS];
    __shared__ uint32_t s_min, s_max;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const B4dGeom &g = p.g;

    long long t = blockIdx.x;
    const int tx = (int)(t % g.tx);
    t /= g.tx;
    const int ty = (int)(t % g.ty);
    t /= g.ty;
    const int tz = (int)(t % g.tz);
    const int vol = (int)(t / g.tz);
    const int iz0 = tz * 4, iy0 = ty * 4, ix0 = tx * 4;
    const int bz = g.refz[iz0] - R_, by = g.refy[iy0] - R_, bx = g.refx[ix0] - R_;
    const uint16_t *__restrict__ uv = p.u + (long long)vol * g.vol_stride;

    if (threadIdx.x == 0) {
        s_min = 0xFFFFFFFFu;
        s_max = 0u;
    }
    __syncthreads();
    {
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        for (int i = threadIdx.x; i < EV; i += WARPS * 32) {
            const int x = i % E, y = (i / E) % E, z = i / (E * E);
            const int gz = bz + z, gy = by + y, gx = bx + x;
            uint32_t v = 0;
            if ((unsigned)gz < (unsigned)g.D && (unsigned)gy < (unsigned)g.H && (unsigned)gx < (unsigned)g.W) {
                v = uv[((long long)gz * g.H + gy) * g.W + gx];
                mn = min(mn, v);
                mx = max(mx, v);
            }
            s_win[i] = (uint16_t)v;
        }
        mn = __reduce_min_sync(B4D_FULL, mn);
        mx = __reduce_max_sync(B4D_FULL, mx);
        if (lane == 0) {
            atomicMin(&s_min, mn);
            atomicMax(&s_max, mx);
        }
    }
    __syncthreads();
    const bool narrow = (s_max - s_min) <= 8191u || s_max < s_min;
    if (!narrow && threadIdx.x == 0 && p.stats) atomicAdd(&p.stats[1], 1ull);

    const int K = p.K;
    const uint32_t tau = p.tau;

    for (int rr = warp; rr < 64; rr += WARPS) {
        const int iz = iz0 + (rr >> 4), iy = iy0 + ((rr >> 2) & 3), ix = ix0 + (rr & 3);
        if (iz >= g.nrz || iy >= g.nry || ix >= g.nrx) continue;  // warp-uniform
        const int oz = g.refz[iz], oy = g.refy[iy], ox = g.refx[ix];
        const int wz0 = oz - R_ - bz, wy0 = oy - R_ - by, wx0 = ox - R_ - bx;  // window origin in the tile
        const uint16_t *refp = s_win + ((wz0 + R_) * E + (wy0 + R_)) * E + (wx0 + R_);
        int ref[B4D_LV];
        if (narrow) {
#pragma unroll
            for (int z = 0; z < 4; ++z)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int x = 0; x < 4; ++x) ref[(z * 4 + y) * 4 + x] = refp[(z * E + y) * E + x];
        }
        // valid dx range of candidates: cx = ox - R_ + j in [0, W-4]
        const int jlo = max(0, R_ - ox), jhi = min(NS - 1, g.W - 4 - ox + R_);

        const long long rlin = (long long)vol * g.refs_per_vol + ((long long)iz * g.nry + iy) * g.nrx + ix;

        bool fallback = false;
        uint32_t prev = 0;      // fallback: last extracted key
        int nsel = 0;           // fallback: keys extracted so far
        uint32_t mykey = B4D_INVALID_KEY;  // fallback: lane k holds the k-th key
        for (;;) {
            uint32_t lmin1 = B4D_INVALID_KEY, lmin2 = B4D_INVALID_KEY;
            uint32_t B = B4D_INVALID_KEY - 1u;
            if (lane == 0) s_cnt[warp] = 0;
            __syncwarp();
#pragma unroll 1
            for (int it = 0; it < ITERS; ++it) {
                const int unit = it * 32 + lane;
                const int dz = unit / NS, dy = unit - dz * NS;
                const int cz = oz - R_ + dz, cy = oy - R_ + dy;
                const bool uvalid = unit < UNITS && cz >= 0 && cz <= g.D - 4 && cy >= 0 && cy <= g.H - 4;
                if (!__any_sync(B4D_FULL, uvalid)) continue;
                uint32_t key[NS];
#pragma unroll
                for (int j = 0; j < NS; ++j) key[j] = B4D_INVALID_KEY;
                if (uvalid) {
                    const uint16_t *base = s_win + ((wz0 + dz) * E + (wy0 + dy)) * E + wx0;
                    if (narrow) {
                        uint32_t acc[NS];
                        ssd_row_u32<NS>(base, ref, acc);
#pragma unroll
                        for (int j = 0; j < NS; ++j) {
                            const bool ok = acc[j] <= tau && j >= jlo && j <= jhi;
                            key[j] = ok ? ((acc[j] << KB) | (uint32_t)(unit * NS + j)) : B4D_INVALID_KEY;
                        }
                    } else {
                        unsigned long long acc[NS];
                        ssd_row_u64<NS>(base, refp, acc);
#pragma unroll
                        for (int j = 0; j < NS; ++j) {
                            const bool ok = acc[j] <= (unsigned long long)tau && j >= jlo && j <= jhi;
                            key[j] = ok ? (((uint32_t)acc[j] << KB) | (uint32_t)(unit * NS + j)) : B4D_INVALID_KEY;
                        }
                    }
                }
                if (!fallback) {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        if (K32) lmin2 = min(lmin2, max(lmin1, key[j]));
                        lmin1 = min(lmin1, key[j]);
                    }
                    const uint32_t b = kth_smallest32(K32 ? lmin2 : lmin1, K32 ? 15 : K - 1, lane);
                    B = min(B, b);
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        if (key[j] <= B) {
                            const int pos = atomicAdd(&s_cnt[warp], 1);
                            if (pos < CAP) s_surv[warp][pos] = key[j];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const uint32_t k = (nsel == 0 || key[j] > prev) ? key[j] : B4D_INVALID_KEY;
                        lmin1 = min(lmin1, k);
                    }
                }
            }
            __syncwarp();
            if (!fallback) {
                const int n = s_cnt[warp];
                if (n <= CAP) {
                    // rank sort of the survivors that are <= the final bound
                    uint32_t mine[CAP / 32];
                    int nf = 0;
#pragma unroll
                    for (int s = 0; s < CAP / 32; ++s) {
                        const int e = s * 32 + lane;
                        uint32_t k = (e < n) ? s_surv[warp][e] : B4D_INVALID_KEY;
                        if (k > B) k = B4D_INVALID_KEY;
                        mine[s] = k;
                        nf += __popc(__ballot_sync(B4D_FULL, k != B4D_INVALID_KEY));
                    }
                    const int ns = min(nf, K);
                    const int kp = ns > 0 ? (1 << (31 - __clz(ns))) : 0;
#pragma unroll
                    for (int s = 0; s < CAP / 32; ++s) {
                        if (s * 32 >= n) break;  // warp-uniform
                        int rank = 0;
                        const uint32_t k = mine[s];
                        for (int e = 0; e < n; ++e) rank += (s_surv[warp][e] < k) ? 1 : 0;
                        if (k != B4D_INVALID_KEY && rank < kp) {
                            p.widx[rlin * K + rank] = (uint16_t)(k & ((1u << KB) - 1u));
                            if (p.ssd_out) p.ssd_out[rlin * K + rank] = k >> KB;
                        }
                    }
                    if (lane == 0) p.cnt[rlin] = (uint8_t)kp;
                    break;
                }
                fallback = true;  // list overflowed: exact but slow path
                if (lane == 0 && p.stats) atomicAdd(&p.stats[0], 1ull);
                continue;
            }
            // fallback round finished: lmin1 = smallest key > prev in this lane
            const uint32_t m = warp_min_u32(lmin1);
            if (m != B4D_INVALID_KEY) {
                if (lane == nsel) mykey = m;
                prev = m;
                ++nsel;
            }
            if (m == B4D_INVALID_KEY || nsel == K) {
                const int kp = nsel > 0 ? (1 << (31 - __clz(nsel))) : 0;
                if (lane < kp) {
                    p.widx[rlin * K + lane] = (uint16_t)(mykey & ((1u << KB) - 1u));
                    if (p.ssd_out) p.ssd_out[rlin * K + lane] = mykey >> KB;
                }
                if (lane == 0) p.cnt[rlin] = (uint8_t)kp;
                break;
            }
        }
        __syncwarp();
    }
}

template <int NS>
void launch_ns(const MatchParams &p, cudaStream_t s) {
    constexpr int E = NS + 12;
    const size_t smem = ((size_t)E * E * E * sizeof(uint16_t) + 15) & ~(size_t)15;
    const long long tiles = (long long)p.g.nvol * p.g.tz * p.g.ty * p.g.tx;
    if (p.K > 16) {
        cudaFuncSetAttribute(k_match<NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_match<NS, true><<<(unsigned)tiles, WARPS * 32, smem, s>>>(p);
    } else {
        cudaFuncSetAttribute(k_match<NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_match<NS, false><<<(unsigned)tiles, WARPS * 32, smem, s>>>(p);
    }
}

}  // namespace

void b4d_launch_match(const MatchParams &p, int Ns, cudaStream_t s) {
    switch (Ns) {
        case 3: launch_ns<3>(p, s); break;
        case 5: launch_ns<5>(p, s); break;
        case 7: launch_ns<7>(p, s); break;
        case 9: launch_ns<9>(p, s); break;
        case 11: launch_ns<11>(p, s); break;
        case 13: launch_ns<13>(p, s); break;
        case 15: launch_ns<15>(p, s); break;
        default: break;
    }
}
