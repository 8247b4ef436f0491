// b4d_misc.cu — the HBM-bound kernels around the two stages (K3/K6 normalise,
// K7 quantize, K8 histogram, dtype preparation) and the issue-rate
// microbenchmarks that give the INT32 / FP32 roofline denominators.
//
// All elementwise kernels are grid-stride over 16-byte vectors with the grid
// sized to a multiple of the SM count; the tail is handled scalar.
#include "b4d_common.cuh"
#include "b4d_tma.cuh"

namespace {

int g_sms = 0;
int sm_count() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}
unsigned grid_for(long long nvec, int threads, int per_sm) {
    long long want = (nvec + threads - 1) / threads;
    long long cap = (long long)sm_count() * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

// ------------------------------------------------------------ u16 -> f32 ----
// Also reduces min / max of the tile (minmax[0] = min, minmax[1] = max): the
// stage-2 matching image is centred in the uint16 range with them.
__global__ void __launch_bounds__(256) k_u16_to_f32(const uint16_t *__restrict__ in, float *__restrict__ out,
                                                    long long n, unsigned *__restrict__ minmax) {
    const long long nv = n >> 3;  // 8 voxels: 16 B in, 32 B out
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned mn = 0xFFFFu, mx = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
        const unsigned e[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                               v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            mn = min(mn, e[k]);
            mx = max(mx, e[k]);
        }
        float4 a, b;
        a.x = (float)e[0];
        a.y = (float)e[1];
        a.z = (float)e[2];
        a.w = (float)e[3];
        b.x = (float)e[4];
        b.y = (float)e[5];
        b.z = (float)e[6];
        b.w = (float)e[7];
        reinterpret_cast<float4 *>(out)[2 * i] = a;
        reinterpret_cast<float4 *>(out)[2 * i + 1] = b;
    }
    for (long long i = (nv << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned e = in[i];
        mn = min(mn, e);
        mx = max(mx, e);
        out[i] = (float)e;
    }
    mn = __reduce_min_sync(B4D_FULL, mn);
    mx = __reduce_max_sync(B4D_FULL, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], mn);
        atomicMax(&minmax[1], mx);
    }
}

// ---------------------------------------------- f32 -> matching image (u16) --
// u = clamp(int(rint((v + cf) * scale)) + ishift, 0, 65535): the integer shift is
// applied after rounding so that rounding never depends on it.
__device__ __forceinline__ uint32_t to_match(float v, float cf, float scale, int ishift) {
    long long q = (long long)__float2ll_rn((v + cf) * scale) + ishift;
    q = min(max(q, 0ll), 65535ll);
    return (uint32_t)q;
}
__global__ void __launch_bounds__(256) k_to_match(const float *__restrict__ in, uint16_t *__restrict__ out,
                                                  long long n, float shift, float scale, int ishift) {
    const long long nv = n >> 3;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 a = reinterpret_cast<const float4 *>(in)[2 * i];
        const float4 b = reinterpret_cast<const float4 *>(in)[2 * i + 1];
        uint4 o;
        o.x = to_match(a.x, shift, scale, ishift) | (to_match(a.y, shift, scale, ishift) << 16);
        o.y = to_match(a.z, shift, scale, ishift) | (to_match(a.w, shift, scale, ishift) << 16);
        o.z = to_match(b.x, shift, scale, ishift) | (to_match(b.y, shift, scale, ishift) << 16);
        o.w = to_match(b.z, shift, scale, ishift) | (to_match(b.w, shift, scale, ishift) << 16);
        reinterpret_cast<uint4 *>(out)[i] = o;
    }
    for (long long i = (nv << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (uint16_t)to_match(in[i], shift, scale, ishift);
}

// ------------------------------------------------------ precompute targets --
// raw = float(u16) - offset[vol] in float32 (data_handling.py:353-354), plus min / max of the counts.
__global__ void __launch_bounds__(256) k_u16_sub_offset(const uint16_t *__restrict__ in, const float *__restrict__ off,
                                                        float *__restrict__ out, long long vol_stride, long long n,
                                                        unsigned *__restrict__ minmax) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned mn = 0xFFFFu, mx = 0u;
    if ((vol_stride & 7) == 0) {
        // 8 voxels per thread and iteration (they belong to one volume): one 16-byte load, two 16-byte stores
        const long long nv = n >> 3, vs8 = vol_stride >> 3;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
            const float o = __ldg(off + i / vs8);
            const unsigned e[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16,
                                   v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mn = min(mn, e[k]);
                mx = max(mx, e[k]);
            }
            __stcs(reinterpret_cast<float4 *>(out) + 2 * i,
                   make_float4(__fsub_rn((float)e[0], o), __fsub_rn((float)e[1], o), __fsub_rn((float)e[2], o),
                               __fsub_rn((float)e[3], o)));
            __stcs(reinterpret_cast<float4 *>(out) + 2 * i + 1,
                   make_float4(__fsub_rn((float)e[4], o), __fsub_rn((float)e[5], o), __fsub_rn((float)e[6], o),
                               __fsub_rn((float)e[7], o)));
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const unsigned e = in[i];
            mn = min(mn, e);
            mx = max(mx, e);
            out[i] = __fsub_rn((float)e, __ldg(off + i / vol_stride));
        }
    }
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], mn);
        atomicMax(&minmax[1], mx);
    }
}
// teacher = clip(x, 0, max_count) in place (data_handling.py:333)
__global__ void __launch_bounds__(256) k_clip(float *__restrict__ x, long long n, float hi) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nv = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 v = reinterpret_cast<float4 *>(x)[i];
        v.x = fminf(fmaxf(v.x, 0.0f), hi);
        v.y = fminf(fmaxf(v.y, 0.0f), hi);
        v.z = fminf(fmaxf(v.z, 0.0f), hi);
        v.w = fminf(fmaxf(v.w, 0.0f), hi);
        reinterpret_cast<float4 *>(x)[i] = v;
    }
    for (long long i = (nv << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        x[i] = fminf(fmaxf(x[i], 0.0f), hi);
}

// ------------------------------------------------------------- normalise ----
// K3 / K6 (weight-map contract): the filter kernels leave, per block origin o, the integer G[o] = sum of the
// group weights qg of the blocks that sit there.  den(v) = sum over d in [0,4)^3 of G[v - d] kf[dz] kf[dy] kf[dx],
// evaluated as three 4-tap passes (x, then y, then z), each an fma chain over d = 0..3 in float64 — the order of
// oracle den_from_weight_map, so the result is identical bit for bit.  Origins outside the volume hold 0, and
// fma(k, 0, acc) == acc, so the tile halo needs no special cases.  out = num / den / qscale, else the fallback.
//
// Algorithmic traffic: int64 numerator 8 B + map 4 B + result 4 B per voxel (the
// fallback is read only where the denominator is zero, which a complete block grid never produces).
// Optional fused outputs: the uint16 matching image of the next stage (K3: saves the separate conversion pass) and
// the quantized uint16 volume (K6 + K7 fused: float32 result never written).
constexpr int WM_TY = 8, WM_TX = 29, WM_NZ = 32, WM_WARPS = 4;
constexpr int WM_EY = WM_TY + 3;
struct NormOut {
    float *out;          // float32 result (may be null when q16 is set)
    uint16_t *match;     // optional: clamp(rint(y * mscale) + ishift, 0, 65535)
    uint16_t *q16;       // optional: K7 of the result
    float mscale;
    int ishift;
    float q_sub, q_add, q_step, q_hi;
    int q_unit, q_trunc;
};
__device__ __forceinline__ uint32_t quant1(float x, float osub, float oadd, float step, float hi, bool unit);
__device__ __forceinline__ uint32_t quant1_trunc(float x, float osub, float oadd, float step, float hi, bool unit);
// One WARP owns an 8 x 29 (y, x) tile and marches along z over WM_NZ planes (+3 planes of run-in); no shared memory,
// no barrier.  Lane i holds column X0 - 3 + i of the uint32 map for the 11 rows Y0 - 3 .. Y0 + 7 (lanes 0-2 are the
// x halo); the x pass takes the three left neighbours by warp shuffles of the uint32 values, the y pass runs on the
// 11 x-convolved rows in registers, the z pass on a rolling window of three earlier xy-convolved values per row.
__global__ void __launch_bounds__(WM_WARPS * 32) k_normalise_wm(const long long *__restrict__ numq,
                                                                const uint32_t *__restrict__ gmap,
                                                                const float *__restrict__ fb, NormOut o, int D, int H,
                                                                int W, int nvol, int z0, int z1, float inv_qscale,
                                                                float kf0, float kf1, float kf2, float kf3) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double k[4] = {(double)kf0, (double)kf1, (double)kf2, (double)kf3};
    const double inv = (double)inv_qscale;
    const int ntx = (W + WM_TX - 1) / WM_TX, nty = (H + WM_TY - 1) / WM_TY, ntz = (z1 - z0 + WM_NZ - 1) / WM_NZ;
    long long t = (long long)blockIdx.x * WM_WARPS + warp;
    if (t >= (long long)nvol * ntz * nty * ntx) return;
    const int ix = (int)(t % ntx);
    t /= ntx;
    const int iy = (int)(t % nty);
    t /= nty;
    const int iz = (int)(t % ntz);
    const long long vol = t / ntz;
    const long long P = (long long)H * W, V = P * D;
    const int X0 = ix * WM_TX, Y0 = iy * WM_TY, Za = z0 + iz * WM_NZ, Zb = min(Za + WM_NZ, z1);
    const uint32_t *g = gmap + vol * V;
    const int gx = X0 - 3 + lane;  // this lane's column
    const bool col_in = (unsigned)gx < (unsigned)W;
    const bool out_lane = lane >= 3 && col_in;
    double h1[WM_TY], h2[WM_TY], h3[WM_TY];  // xy-convolved map of the three previous planes
#pragma unroll
    for (int r = 0; r < WM_TY; ++r) h1[r] = h2[r] = h3[r] = 0.0;
    // Memory latency is the limit of this kernel (ncu: long-scoreboard stalls, DRAM at 17 %), so every load is
    // issued well ahead of its use: the map values of plane z + 1 and the numerators of plane z before plane z is
    // convolved.
    uint32_t gn[WM_EY];
    auto load_map = [&](int z) {
#pragma unroll
        for (int r = 0; r < WM_EY; ++r) {
            const int yy = Y0 - 3 + r;
            gn[r] = (z >= 0 && col_in && (unsigned)yy < (unsigned)H) ? __ldg(g + (long long)z * P + (long long)yy * W + gx) : 0u;
        }
    };
    load_map(Za - 3);
    for (int z = Za - 3; z < Zb; ++z) {
        uint32_t gv[WM_EY];
#pragma unroll
        for (int r = 0; r < WM_EY; ++r) gv[r] = gn[r];
        if (z + 1 < Zb) load_map(z + 1);
        const bool out_z = z >= Za && out_lane;
        const long long a0 = vol * V + (long long)z * P + (long long)Y0 * W + gx;
        long long nq[WM_TY];
#pragma unroll
        for (int r = 0; r < WM_TY; ++r) nq[r] = (out_z && Y0 + r < H) ? __ldcs(numq + a0 + (long long)r * W) : 0ll;
        double xr[WM_EY];  // x pass: sum over d of k[d] G[x - d]
#pragma unroll
        for (int r = 0; r < WM_EY; ++r) {
            double acc = fma(k[0], (double)gv[r], 0.0);
#pragma unroll
            for (int d = 1; d < 4; ++d) acc = fma(k[d], (double)__shfl_up_sync(B4D_FULL, gv[r], d), acc);
            xr[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < WM_TY; ++r) {
            double w0 = 0.0;  // y pass
#pragma unroll
            for (int d = 0; d < 4; ++d) w0 = fma(k[d], xr[r + 3 - d], w0);
            if (out_z && Y0 + r < H) {
                double den = fma(k[0], w0, 0.0);
                den = fma(k[1], h1[r], den);
                den = fma(k[2], h2[r], den);
                den = fma(k[3], h3[r], den);
                const long long a = a0 + (long long)r * W;
                const float y = den > 0.0 ? (float)(((double)nq[r] / den) * inv) : fb[a];
                if (o.out) o.out[a] = y;
                if (o.match) o.match[a] = (uint16_t)to_match(y, 0.0f, o.mscale, o.ishift);
                if (o.q16)
                    o.q16[a] = (uint16_t)(o.q_trunc ? quant1_trunc(y, o.q_sub, o.q_add, o.q_step, o.q_hi, o.q_unit != 0)
                                                    : quant1(y, o.q_sub, o.q_add, o.q_step, o.q_hi, o.q_unit != 0));
            }
            h3[r] = h2[r];
            h2[r] = h1[r];
            h1[r] = w0;
        }
    }
}

// The same arithmetic with the inputs brought in by TMA (the production path when W % 4 == 0).  A CTA of 4 warps owns
// 8 x 112 (y, x) voxels and marches along z; per plane two box loads (the map: 120 x 11 uint32 from (x0 - 4, Y0 - 3),
// zero fill outside the volume = "no block there"; the numerators: 112 x 8 int64) land in a ring of WMT_NST stages,
// full / empty mbarriers per stage, thread 0 refills the stage of the previous plane.  ~50 KB of loads per CTA are in
// flight without costing registers; the warps only read shared memory.  Lane i of warp w holds map column
// x0 + 28 w - 3 + i (lanes 3-30 produce outputs).
constexpr int WMT_WARPS = 4, WMT_WX = 28, WMT_TX = WMT_WARPS * WMT_WX, WMT_GW = 120, WMT_NST = 4;
constexpr uint32_t WMT_G_BYTES = WMT_GW * WM_EY * 4, WMT_N_BYTES = WMT_TX * WM_TY * 8;
constexpr int WMT_G_STAGE = ((int)WMT_G_BYTES + 127) / 128 * 128;
constexpr int WMT_STAGE = WMT_G_STAGE + (int)WMT_N_BYTES;
static_assert(WMT_TX % 4 == 0 && WMT_GW % 4 == 0 && WMT_GW >= WMT_TX + 4 && WMT_N_BYTES % 128 == 0, "TMA boxes of the normalise kernel");
__global__ void __launch_bounds__(WMT_WARPS * 32, 4) k_normalise_wm_tma(const __grid_constant__ CUtensorMap gmap_t,
                                                                       const __grid_constant__ CUtensorMap numq_t,
                                                                       const float *__restrict__ fb, NormOut o, int D, int H,
                                                                       int W, int nvol, int z0, int z1, float inv_qscale,
                                                                       float kf0, float kf1, float kf2, float kf3) {
    extern __shared__ __align__(128) unsigned char s_buf[];
    __shared__ __align__(8) unsigned long long s_full[WMT_NST], s_empty[WMT_NST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double k[4] = {(double)kf0, (double)kf1, (double)kf2, (double)kf3};
    const double inv = (double)inv_qscale;
    const int ntx = (W + WMT_TX - 1) / WMT_TX, nty = (H + WM_TY - 1) / WM_TY, ntz = (z1 - z0 + WM_NZ - 1) / WM_NZ;
    long long t = blockIdx.x;
    const int ix = (int)(t % ntx);
    t /= ntx;
    const int iy = (int)(t % nty);
    t /= nty;
    const int iz = (int)(t % ntz);
    const int vol = (int)(t / ntz);
    const long long P = (long long)H * W, V = P * D;
    const int x0 = ix * WMT_TX, Y0 = iy * WM_TY, Za = z0 + iz * WM_NZ, Zb = min(Za + WM_NZ, z1);
    const int nplanes = Zb - (Za - 3);
    const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(s_buf);
    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(s_full), empty0 = (uint32_t)__cvta_generic_to_shared(s_empty);
    if (threadIdx.x == 0) {
        for (int i = 0; i < WMT_NST; ++i) {
            mbar_init(full0 + 8u * i, 1);
            mbar_init(empty0 + 8u * i, WMT_WARPS);
        }
    }
    __syncthreads();
    auto issue = [&](int i) {  // plane Za - 3 + i into stage i % WMT_NST (thread 0)
        const int st = i % WMT_NST, z = Za - 3 + i;
        const uint32_t dst = buf0 + (uint32_t)(st * WMT_STAGE), bar = full0 + 8u * st;
        mbar_expect_tx(bar, WMT_G_BYTES + (i >= 3 ? WMT_N_BYTES : 0u));
        tma_load_4d(dst, &gmap_t, bar, x0 - 4, Y0 - 3, z, vol);
        if (i >= 3) tma_load_4d(dst + WMT_G_STAGE, &numq_t, bar, x0, Y0, z, vol);  // run-in planes need no numerators
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < min(WMT_NST, nplanes); ++i) issue(i);
    const int gx = x0 + warp * WMT_WX - 3 + lane;  // this lane's column
    const bool out_lane = lane >= 3 && lane < 3 + WMT_WX && gx < W;
    double h1[WM_TY], h2[WM_TY], h3[WM_TY];  // xy-convolved map of the three previous planes
#pragma unroll
    for (int r = 0; r < WM_TY; ++r) h1[r] = h2[r] = h3[r] = 0.0;
    for (int i = 0; i < nplanes; ++i) {
        const int st = i % WMT_NST, z = Za - 3 + i;
        mbar_wait(full0 + 8u * st, (uint32_t)((i / WMT_NST) & 1));
        const unsigned char *sb = s_buf + st * WMT_STAGE;
        const uint32_t *sg = reinterpret_cast<const uint32_t *>(sb) + warp * WMT_WX + 1 + lane;  // column gx - (x0 - 4)
        const long long *sn = reinterpret_cast<const long long *>(sb + WMT_G_STAGE) + warp * WMT_WX + (lane - 3);
        uint32_t gv[WM_EY];
#pragma unroll
        for (int r = 0; r < WM_EY; ++r) gv[r] = sg[r * WMT_GW];
        const bool out_z = i >= 3 && out_lane;
        long long nq[WM_TY];
#pragma unroll
        for (int r = 0; r < WM_TY; ++r) nq[r] = out_z ? sn[r * WMT_TX] : 0ll;
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * st);
        if (threadIdx.x == 0 && i >= 1 && i - 1 + WMT_NST < nplanes) {  // refill the stage of the previous plane
            mbar_wait(empty0 + 8u * ((i - 1) % WMT_NST), (uint32_t)(((i - 1) / WMT_NST) & 1));
            issue(i - 1 + WMT_NST);
        }
        double xr[WM_EY];  // x pass: sum over d of k[d] G[x - d]
#pragma unroll
        for (int r = 0; r < WM_EY; ++r) {
            double acc = fma(k[0], (double)gv[r], 0.0);
#pragma unroll
            for (int d = 1; d < 4; ++d) acc = fma(k[d], (double)__shfl_up_sync(B4D_FULL, gv[r], d), acc);
            xr[r] = acc;
        }
        const long long a0 = (long long)vol * V + (long long)z * P + (long long)Y0 * W + gx;
#pragma unroll
        for (int r = 0; r < WM_TY; ++r) {
            double w0 = 0.0;  // y pass
#pragma unroll
            for (int d = 0; d < 4; ++d) w0 = fma(k[d], xr[r + 3 - d], w0);
            if (out_z && Y0 + r < H) {
                double den = fma(k[0], w0, 0.0);
                den = fma(k[1], h1[r], den);
                den = fma(k[2], h2[r], den);
                den = fma(k[3], h3[r], den);
                const long long a = a0 + (long long)r * W;
                const float y = den > 0.0 ? (float)(((double)nq[r] / den) * inv) : fb[a];
                if (o.out) o.out[a] = y;
                if (o.match) o.match[a] = (uint16_t)to_match(y, 0.0f, o.mscale, o.ishift);
                if (o.q16)
                    o.q16[a] = (uint16_t)(o.q_trunc ? quant1_trunc(y, o.q_sub, o.q_add, o.q_step, o.q_hi, o.q_unit != 0)
                                                    : quant1(y, o.q_sub, o.q_add, o.q_step, o.q_hi, o.q_unit != 0));
            }
            h3[r] = h2[r];
            h2[r] = h1[r];
            h1[r] = w0;
        }
    }
}

// -------------------------------------------------------------- quantize ----
// K7: q = rint(clip((x - offset_sub + offset_add) / step, 0, 65535/step)).
// float32, clip BEFORE round, round-half-to-even: transforms.py:403-411.
__device__ __forceinline__ uint32_t quant1(float x, float osub, float oadd, float step, float hi, bool unit) {
    float v = __fadd_rn(__fsub_rn(x, osub), oadd);
    if (!unit) v = __fdiv_rn(v, step);
    v = fminf(fmaxf(v, 0.0f), hi);
    return (uint32_t)__float2int_rn(v);
}
// truncating variant: np.maximum(x, 0).astype(int) (evaluate.py:202) followed by the uint16 cast of
// compute_cratio (utils/img_util.py:420-423): toward zero, no upper clip, int64 -> uint16 wraps modulo 2^16
__device__ __forceinline__ uint32_t quant1_trunc(float x, float osub, float oadd, float step, float hi, bool unit) {
    (void)hi;
    float v = __fadd_rn(__fsub_rn(x, osub), oadd);
    if (!unit) v = __fdiv_rn(v, step);
    v = fmaxf(v, 0.0f);  // NaN -> 0
    return (uint32_t)(__float2ll_rz(v) & 0xFFFFll);
}
template <bool TRUNC>
__device__ __forceinline__ uint32_t quantq(float x, float osub, float oadd, float step, float hi, bool unit) {
    return TRUNC ? quant1_trunc(x, osub, oadd, step, hi, unit) : quant1(x, osub, oadd, step, hi, unit);
}
template <bool TRUNC>
__global__ void __launch_bounds__(256) k_quantize(const float *__restrict__ in, uint16_t *__restrict__ out,
                                                  long long n, float osub, float oadd, float step) {
    const bool unit = (step == 1.0f);
    const float hi = __fdiv_rn(65535.0f, step);
    // 16 voxels per thread and iteration: four 16-byte loads in flight, two 16-byte stores
    const long long nv = n >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldcs(reinterpret_cast<const float4 *>(in) + 4 * i + k);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 a = v[2 * k], b = v[2 * k + 1];
            uint4 o;
            o.x = quantq<TRUNC>(a.x, osub, oadd, step, hi, unit) | (quantq<TRUNC>(a.y, osub, oadd, step, hi, unit) << 16);
            o.y = quantq<TRUNC>(a.z, osub, oadd, step, hi, unit) | (quantq<TRUNC>(a.w, osub, oadd, step, hi, unit) << 16);
            o.z = quantq<TRUNC>(b.x, osub, oadd, step, hi, unit) | (quantq<TRUNC>(b.y, osub, oadd, step, hi, unit) << 16);
            o.w = quantq<TRUNC>(b.z, osub, oadd, step, hi, unit) | (quantq<TRUNC>(b.w, osub, oadd, step, hi, unit) << 16);
            __stcs(reinterpret_cast<uint4 *>(out) + 2 * i + k, o);
        }
    }
    for (long long i = (nv << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (uint16_t)quantq<TRUNC>(in[i], osub, oadd, step, hi, unit);
}

// --------------------------------------------------- chunk byte shuffle ----
// K9 (first step): the chunking of compute_cratio (img_util.py:401-441) — a C-order
// grid of cz x cy x cx pieces, ragged at the far faces, each piece made contiguous — followed by the
// byte shuffle Blosc SHUFFLE applies to 2-byte items: all low bytes of the piece, then all high bytes.
// Piece (iz, iy, ix) starts at element z0*H*W + dz*(y0*W + dy*x0) of the output (the pieces before it
// in C order hold exactly that many voxels).  Also counts the byte values of both planes per piece
// (hist[piece][plane][256]); a warp whose 128 bytes are all equal (the high plane of background) adds once.
// Measured on B200: 32 lane-private replicas of the counters (no intra-warp collisions) were SLOWER than
// one copy — the kernel is bound by the memory pipeline, not by the shared-memory atomics.
__device__ __forceinline__ void red_shared_add(uint32_t *p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void hist_add4(uint32_t *sh, uint32_t b4, bool valid, int nvalid) {
    const uint32_t b0 = __shfl_sync(B4D_FULL, b4, 0);
    const bool uni = __all_sync(B4D_FULL, !valid || b4 == b0) && b0 == (b0 & 0xFFu) * 0x01010101u;
    if (uni) {
        if ((threadIdx.x & 31) == 0) red_shared_add(&sh[b0 & 0xFFu], 4u * nvalid);
    } else if (valid) {
        if (b4 == (b4 & 0xFFu) * 0x01010101u) {
            red_shared_add(&sh[b4 & 0xFFu], 4u);
        } else {
            red_shared_add(&sh[b4 & 0xFFu], 1u);
            red_shared_add(&sh[(b4 >> 8) & 0xFFu], 1u);
            red_shared_add(&sh[(b4 >> 16) & 0xFFu], 1u);
            red_shared_add(&sh[b4 >> 24], 1u);
        }
    }
}
// One piece whose rows are whole groups of V voxels (V = 4: 8-byte loads, V = 8: 16-byte loads).  Group
// t = row * q + k holds elements [V t, V t + V) of the piece; (row, k) and (z, y) advance incrementally with
// carries (no division in the loop) and UN loads are in flight per thread.
template <int V>
__device__ __forceinline__ void shuffle_piece_vec(const uint16_t *__restrict__ src0, int H, int W, int dz, int dy,
                                                  int dx, uint8_t *lo, uint8_t *hi, uint32_t *sh_lo, uint32_t *sh_hi,
                                                  bool hist) {
    constexpr int UN = 4, WORDS = V / 4;  // 32-bit words of low (and of high) bytes per group
    const int q = dx / V, total = dz * dy * q;
    const int dk = 256 % q, dr = 256 / q;
    int k = (int)threadIdx.x % q, row = (int)threadIdx.x / q;
    int z = row / dy, y = row - z * dy;
    for (int base = 0; base < total; base += 256 * UN) {
        uint32_t v[UN][2 * WORDS];
        bool valid[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            valid[u] = base + u * 256 + (int)threadIdx.x < total;
#pragma unroll
            for (int w = 0; w < 2 * WORDS; ++w) v[u][w] = 0u;
            if (valid[u]) {
                const uint16_t *p = src0 + ((long long)z * H + y) * W + V * k;
                if (V == 8) {
                    const uint4 t4 = __ldcs(reinterpret_cast<const uint4 *>(p));
                    v[u][0] = t4.x, v[u][1] = t4.y, v[u][2 * WORDS - 2] = t4.z, v[u][2 * WORDS - 1] = t4.w;
                } else {
                    const uint2 t2 = __ldcs(reinterpret_cast<const uint2 *>(p));
                    v[u][0] = t2.x, v[u][1] = t2.y;
                }
            }
            k += dk;
            int step = dr;
            if (k >= q) {
                k -= q;
                ++step;
            }
            y += step;
            if (y >= dy) {
                if (step <= dy) {
                    y -= dy;
                    ++z;
                } else {
                    z += y / dy;
                    y %= dy;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int t = base + u * 256 + (int)threadIdx.x;
            uint32_t lo4[WORDS], hi4[WORDS];
#pragma unroll
            for (int w = 0; w < WORDS; ++w) {
                lo4[w] = __byte_perm(v[u][2 * w], v[u][2 * w + 1], 0x6420);
                hi4[w] = __byte_perm(v[u][2 * w], v[u][2 * w + 1], 0x7531);
            }
            if (lo && valid[u]) {
                if (V == 8) {
                    __stcs(reinterpret_cast<uint2 *>(lo) + t, make_uint2(lo4[0], lo4[WORDS - 1]));
                    __stcs(reinterpret_cast<uint2 *>(hi) + t, make_uint2(hi4[0], hi4[WORDS - 1]));
                } else {
                    __stcs(reinterpret_cast<uint32_t *>(lo) + t, lo4[0]);
                    __stcs(reinterpret_cast<uint32_t *>(hi) + t, hi4[0]);
                }
            }
            if (hist) {
                const int nvalid = __popc(__ballot_sync(B4D_FULL, valid[u]));
                if (nvalid) {
#pragma unroll
                    for (int w = 0; w < WORDS; ++w) {
                        hist_add4(sh_lo, lo4[w], valid[u], nvalid);
                        hist_add4(sh_hi, hi4[w], valid[u], nvalid);
                    }
                }
            }
        }
    }
}
__global__ void __launch_bounds__(256) k_chunk_shuffle(const uint16_t *__restrict__ in, int D, int H, int W, int cz,
                                                       int cy, int cx, uint8_t *__restrict__ out,
                                                       uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh_lo[256], sh_hi[256];
    const int nz = (D + cz - 1) / cz, ny = (H + cy - 1) / cy, nx = (W + cx - 1) / cx;
    const long long nchunks = (long long)nz * ny * nx;
    const uintptr_t ai = reinterpret_cast<uintptr_t>(in), ao = reinterpret_cast<uintptr_t>(out);
    const bool al8 = (W & 7) == 0 && (ai & 15) == 0 && (ao & 7) == 0;
    const bool al4 = (W & 3) == 0 && (ai & 7) == 0 && (ao & 3) == 0;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (hist) {
            sh_lo[threadIdx.x] = 0u;
            sh_hi[threadIdx.x] = 0u;
            __syncthreads();
        }
        const int ix = (int)(c % nx), iy = (int)((c / nx) % ny), iz = (int)(c / ((long long)nx * ny));
        const int z0 = iz * cz, y0 = iy * cy, x0 = ix * cx;
        const int dz = min(cz, D - z0), dy = min(cy, H - y0), dx = min(cx, W - x0);
        const long long eoff = (long long)z0 * H * W + (long long)dz * ((long long)y0 * W + (long long)dy * x0);
        const int ne = dz * dy * dx;
        uint8_t *lo = out ? out + 2 * eoff : nullptr;
        uint8_t *hi = out ? lo + ne : nullptr;
        const uint16_t *src0 = in + ((long long)z0 * H + y0) * W + x0;
        if (al8 && (dx & 7) == 0 && (x0 & 7) == 0) {
            shuffle_piece_vec<8>(src0, H, W, dz, dy, dx, lo, hi, sh_lo, sh_hi, hist != nullptr);
        } else if (al4 && (dx & 3) == 0 && (x0 & 3) == 0) {
            shuffle_piece_vec<4>(src0, H, W, dz, dy, dx, lo, hi, sh_lo, sh_hi, hist != nullptr);
        } else {
            for (int e = threadIdx.x; e < ne; e += 256) {
                const int row = e / dx;
                const int x = e - row * dx, z = row / dy, y = row - z * dy;
                const uint32_t v = src0[((long long)z * H + y) * W + x];
                if (out) {
                    lo[e] = (uint8_t)(v & 0xFFu);
                    hi[e] = (uint8_t)(v >> 8);
                }
                if (hist) {
                    red_shared_add(&sh_lo[v & 0xFFu], 1u);
                    red_shared_add(&sh_hi[v >> 8], 1u);
                }
            }
        }
        if (hist) {
            __syncthreads();
            hist[c * 512 + threadIdx.x] = sh_lo[threadIdx.x];
            hist[c * 512 + 256 + threadIdx.x] = sh_hi[threadIdx.x];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------- foreground mask ----
// make_foreground_mask (metrics.py:58-60) once the robust threshold is known: mask = raw > thr with
// raw = float32(u16) - offset (data_handling.py:353-354), then `dilate` iterations of binary dilation
// with the 6-neighbour structuring element and border value 0 — i.e. the L1 ball of radius `dilate`,
// clipped at the patch faces.  One thread per voxel; thr / off are per patch.
__global__ void __launch_bounds__(256) k_fg_mask(const uint16_t *__restrict__ in, const float *__restrict__ off,
                                                 const float *__restrict__ thr, int D, int H, int W, long long n,
                                                 int dilate, uint8_t *__restrict__ out) {
    const long long V = (long long)D * H * W;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long pi = i / V;
        const int r = (int)(i - pi * V);
        const int z = r / (H * W), y = (r / W) % H, x = r % W;
        const float o = __ldg(off + pi), t = __ldg(thr + pi);
        const uint16_t *vol = in + pi * V;
        bool m = false;
        for (int dz = -dilate; dz <= dilate && !m; ++dz) {
            const int zz = z + dz;
            if (zz < 0 || zz >= D) continue;
            const int ry = dilate - abs(dz);
            for (int dy = -ry; dy <= ry && !m; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                const int rx = ry - abs(dy);
                const int x0 = max(0, x - rx), x1 = min(W - 1, x + rx);
                const uint16_t *row = vol + ((long long)zz * H + yy) * W;
                for (int xx = x0; xx <= x1; ++xx) m = m || (__fsub_rn((float)row[xx], o) > t);
            }
        }
        out[i] = m ? 1 : 0;
    }
}

// ------------------------------------------------------------- histogram ----
// K8: exact 65536-bin histogram of a uint16 tile.  Bins below HOT live in a
// per-CTA shared-memory histogram (ExaSPIM background sits there), the rest go
// straight to global atomics; the shared part is flushed once per CTA.
constexpr int HOT = 12288;
__global__ void __launch_bounds__(512) k_hist(const uint16_t *__restrict__ in, long long n,
                                              unsigned long long *__restrict__ hist) {
    __shared__ unsigned int sh[HOT];
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto add = [&](uint32_t v) {
        if (v < HOT) atomicAdd(&sh[v], 1u);
        else atomicAdd(&hist[v], 1ull);
    };
    // elements before the first 16-byte boundary (a patch inside a batch may start anywhere)
    const long long head = min(n, (long long)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(in) & 15u)) & 15u) >> 1));
    if (blockIdx.x == 0 && threadIdx.x < head) add(in[threadIdx.x]);
    in += head;
    n -= head;
    const long long nv = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
        add(v.x & 0xFFFFu);
        add(v.x >> 16);
        add(v.y & 0xFFFFu);
        add(v.y >> 16);
        add(v.z & 0xFFFFu);
        add(v.z >> 16);
        add(v.w & 0xFFFFu);
        add(v.w >> 16);
    }
    for (long long i = (nv << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) add(in[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < HOT; i += blockDim.x) {
        const unsigned int c = sh[i];
        if (c) atomicAdd(&hist[i], (unsigned long long)c);
    }
}

// -------------------------------------------- float32 analysis (match map) --
// per-block partials: {max |frac dev|, min rint, max rint, min z, max z, 1 if any value is NaN or infinite}
constexpr int AN_BLOCKS = 592;
__global__ void __launch_bounds__(256) k_analyze(const float *__restrict__ in, long long n, double c,
                                                 double *__restrict__ partial) {
    double dev = 0.0, lo = 1e300, hi = -1e300, zlo = 1e300, zhi = -1e300, nonfinite = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float zf32 = in[i];
        if (!isfinite(zf32)) nonfinite = 1.0;  // fmin / fmax drop NaNs: count them explicitly
        const double z = (double)zf32;
        const double v = z + c, rv = rint(v);
        dev = fmax(dev, fabs(v - rv));
        lo = fmin(lo, rv);
        hi = fmax(hi, rv);
        zlo = fmin(zlo, z);
        zhi = fmax(zhi, z);
    }
    __shared__ double sh[6][8];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        nonfinite = fmax(nonfinite, __shfl_xor_sync(B4D_FULL, nonfinite, m));
        dev = fmax(dev, __shfl_xor_sync(B4D_FULL, dev, m));
        lo = fmin(lo, __shfl_xor_sync(B4D_FULL, lo, m));
        hi = fmax(hi, __shfl_xor_sync(B4D_FULL, hi, m));
        zlo = fmin(zlo, __shfl_xor_sync(B4D_FULL, zlo, m));
        zhi = fmax(zhi, __shfl_xor_sync(B4D_FULL, zhi, m));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sh[0][warp] = dev;
        sh[1][warp] = lo;
        sh[2][warp] = hi;
        sh[3][warp] = zlo;
        sh[4][warp] = zhi;
        sh[5][warp] = nonfinite;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            sh[0][0] = fmax(sh[0][0], sh[0][w]);
            sh[1][0] = fmin(sh[1][0], sh[1][w]);
            sh[2][0] = fmax(sh[2][0], sh[2][w]);
            sh[3][0] = fmin(sh[3][0], sh[3][w]);
            sh[4][0] = fmax(sh[4][0], sh[4][w]);
            sh[5][0] = fmax(sh[5][0], sh[5][w]);
        }
        for (int q = 0; q < 6; ++q) partial[blockIdx.x * 6 + q] = sh[q][0];
    }
}

// ------------------------------------------------- issue-rate microbenchmarks
// which: 0 IMAD, 1 IADD3, 2 dependent (sub, mad) pairs, 3 FFMA.  8 independent
// chains per thread; 1024 threads per CTA, 2 CTAs per SM.
template <int WHICH>
__global__ void __launch_bounds__(1024) k_pipe(int iters, unsigned *sink) {
    unsigned a[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = threadIdx.x * 2654435761u + i;
        f[i] = (float)(threadIdx.x + i) * 1e-3f;
    }
    const unsigned m = threadIdx.x | 1u, c = blockIdx.x + 3u;
    const float fm = 1.0f + (float)threadIdx.x * 1e-9f, fc = (float)blockIdx.x * 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (WHICH == 0) a[i] = a[i] * m + c;
                if (WHICH == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
                if (WHICH == 2) {
                    const int d = (int)a[i] - (int)c;       // sub
                    a[i] = (unsigned)(d * d) + a[i];        // mad
                }
                if (WHICH == 3) f[i] = __fmaf_rn(f[i], fm, fc);
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ __float_as_uint(f[i]);
    if (r == 0x12345678u) sink[0] = r;
}

}  // namespace

void b4d_launch_u16_to_f32(const uint16_t *in, float *out, long long n, unsigned *minmax, cudaStream_t s) {
    k_u16_to_f32<<<grid_for(n >> 3, 256, 8), 256, 0, s>>>(in, out, n, minmax);
}
void b4d_launch_u16_sub_offset(const uint16_t *in, const float *off, float *out, long long vol_stride, long long n,
                               unsigned *minmax, cudaStream_t s) {
    k_u16_sub_offset<<<grid_for((vol_stride & 7) ? n : (n >> 3), 256, 8), 256, 0, s>>>(in, off, out, vol_stride, n, minmax);
}
__global__ void __launch_bounds__(256) k_add_scalar(float *__restrict__ x, long long n, float c) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = __fadd_rn(x[i], c);
}
void b4d_launch_add_scalar(float *x, long long n, float c, cudaStream_t s) {
    k_add_scalar<<<grid_for(n, 256, 8), 256, 0, s>>>(x, n, c);
}
void b4d_launch_clip(float *x, long long n, float hi, cudaStream_t s) {
    k_clip<<<grid_for(n >> 2, 256, 8), 256, 0, s>>>(x, n, hi);
}
void b4d_launch_to_match(const float *in, uint16_t *out, long long n, float cf, float scale, int ishift,
                         cudaStream_t s) {
    k_to_match<<<grid_for(n >> 3, 256, 8), 256, 0, s>>>(in, out, n, cf, scale, ishift);
}
static void launch_norm(const long long *numq, const uint32_t *gmap, const float *fallback, const NormOut &o, int D,
                        int H, int W, int nvol, int z0, int z1, float inv_qscale, const float kf[4], cudaStream_t s) {
    if (z1 <= z0) return;
    const long long ntz = (z1 - z0 + WM_NZ - 1) / WM_NZ, nty = (H + WM_TY - 1) / WM_TY;
    CUtensorMap mg, mn;
    if ((W & 3) == 0 && !getenv("B4D_NO_TMA") &&
        make_map_4d(&mg, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, gmap, W, H, D, nvol, WMT_GW, WM_EY, 1) &&
        make_map_4d(&mn, CU_TENSOR_MAP_DATA_TYPE_INT64, 8, numq, W, H, D, nvol, WMT_TX, WM_TY, 1)) {
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(k_normalise_wm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, WMT_NST * WMT_STAGE);
            attr = true;
        }
        const long long blocks = (long long)nvol * ntz * nty * ((W + WMT_TX - 1) / WMT_TX);
        k_normalise_wm_tma<<<(unsigned)blocks, WMT_WARPS * 32, WMT_NST * WMT_STAGE, s>>>(
            mg, mn, fallback, o, D, H, W, nvol, z0, z1, inv_qscale, kf[0], kf[1], kf[2], kf[3]);
        return;
    }
    const long long tiles = (long long)nvol * ntz * nty * ((W + WM_TX - 1) / WM_TX);
    k_normalise_wm<<<(unsigned)((tiles + WM_WARPS - 1) / WM_WARPS), WM_WARPS * 32, 0, s>>>(
        numq, gmap, fallback, o, D, H, W, nvol, z0, z1, inv_qscale, kf[0], kf[1], kf[2], kf[3]);
}
void b4d_launch_normalise_wm(const long long *numq, const uint32_t *gmap, const float *fallback, float *out, int D,
                             int H, int W, int nvol, int z0, int z1, float inv_qscale, const float kf[4],
                             cudaStream_t s) {
    NormOut o{};
    o.out = out;
    launch_norm(numq, gmap, fallback, o, D, H, W, nvol, z0, z1, inv_qscale, kf, s);
}
// K6 + K7 fused: the quantized uint16 volume is written directly (float32 result never stored)
void b4d_launch_normalise_q16(const long long *numq, const uint32_t *gmap, const float *fallback, uint16_t *q16, int D,
                              int H, int W, int nvol, int z0, int z1, float inv_qscale, const float kf[4],
                              float offset_sub, float offset_add, float step, int trunc, cudaStream_t s) {
    NormOut o{};
    o.q16 = q16;
    o.q_sub = offset_sub;
    o.q_add = offset_add;
    o.q_step = step;
    o.q_hi = 65535.0f / step;  // float32 division, as k_quantize's __fdiv_rn(65535.0f, step)
    o.q_unit = step == 1.0f;
    o.q_trunc = trunc;
    launch_norm(numq, gmap, fallback, o, D, H, W, nvol, z0, z1, inv_qscale, kf, s);
}
void b4d_launch_normalise_match(const long long *numq, const uint32_t *gmap, const float *fallback, float *out,
                                uint16_t *match, float mscale, int ishift, int D, int H, int W, int nvol,
                                float inv_qscale, const float kf[4], cudaStream_t s) {
    NormOut o{};
    o.out = out;
    o.match = match;
    o.mscale = mscale;
    o.ishift = ishift;
    launch_norm(numq, gmap, fallback, o, D, H, W, nvol, 0, D, inv_qscale, kf, s);
}
void b4d_launch_quantize(const float *in, uint16_t *out, long long n, float offset_sub, float offset_add,
                         float step, cudaStream_t s) {
    k_quantize<false><<<grid_for(n >> 4, 256, 8), 256, 0, s>>>(in, out, n, offset_sub, offset_add, step);
}
void b4d_launch_quantize_trunc(const float *in, uint16_t *out, long long n, float offset_sub, float offset_add,
                               float step, cudaStream_t s) {
    k_quantize<true><<<grid_for(n >> 4, 256, 8), 256, 0, s>>>(in, out, n, offset_sub, offset_add, step);
}
void b4d_launch_chunk_shuffle(const uint16_t *in, int D, int H, int W, int cz, int cy, int cx, uint8_t *out,
                              uint32_t *hist, cudaStream_t s) {
    const long long nchunks = (long long)((D + cz - 1) / cz) * ((H + cy - 1) / cy) * ((W + cx - 1) / cx);
    k_chunk_shuffle<<<grid_for(nchunks * 256, 256, 8), 256, 0, s>>>(in, D, H, W, cz, cy, cx, out, hist);
}
void b4d_launch_fg_mask(const uint16_t *in, const float *off, const float *thr, int D, int H, int W, long long n,
                        int dilate, uint8_t *out, cudaStream_t s) {
    k_fg_mask<<<grid_for(n, 256, 8), 256, 0, s>>>(in, off, thr, D, H, W, n, dilate, out);
}
void b4d_launch_hist(const uint16_t *in, long long n, unsigned long long *hist, cudaStream_t s) {
    k_hist<<<grid_for(n >> 3, 512, 2), 512, 0, s>>>(in, n, hist);
}
int b4d_analyze_blocks() { return AN_BLOCKS; }
void b4d_launch_analyze(const float *in, long long n, double c, double *partial, cudaStream_t s) {
    k_analyze<<<AN_BLOCKS, 256, 0, s>>>(in, n, c, partial);
}
double b4d_launch_pipe_bench(int which, int iters, unsigned *sink, cudaStream_t s) {
    const int blocks = sm_count() * 2;
    switch (which) {
        case 0: k_pipe<0><<<blocks, 1024, 0, s>>>(iters, sink); break;
        case 1: k_pipe<1><<<blocks, 1024, 0, s>>>(iters, sink); break;
        case 2: k_pipe<2><<<blocks, 1024, 0, s>>>(iters, sink); break;
        default: k_pipe<3><<<blocks, 1024, 0, s>>>(iters, sink); break;
    }
    const double per_iter = (which == 2) ? 2.0 : 1.0;
    return (double)blocks * 1024.0 * (double)iters * 16.0 * 8.0 * per_iter;
}
