// Probe: cp.reduce.async.bulk.tensor (.add) of int64 rows from shared memory into a 4-D int64 tensor — the flush of
// the filter kernels' numerators without LSU atomics.  Checks: sums of many CTAs adding overlapping rows, negative
// and out-of-range start coordinates (elements outside the tensor must be skipped), odd/even inner start.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int BOXW = 24;
__global__ void k(const __grid_constant__ CUtensorMap map, long long *gbase, int mode, int W, int H, int D, int reps) {
    __shared__ alignas(128) long long stg[8][BOXW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // CTA b adds, to row (y, z) = (b % H', ...) starting at x0 = 12 * (b % 7) - 6 (even, may be negative), the values
    // value(x) = (x + 1) * (b + 1) for its 24 columns; repeated reps times with slot reuse.
    const int b = blockIdx.x;
    const int x0 = 12 * (b % 9) - 6, y = (b * 5) % (H + 2) - 1, z = (b * 3) % D;   // y may be -1 or H: skipped rows
    for (int r = 0; r < reps; ++r) {
        const int slot = r % 8;
        if (lane == 0 && warp == 0 && r >= 8) asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
        __syncwarp();
        if (warp == 0) {
            if (lane < BOXW) stg[slot][lane] = (long long)(x0 + lane + 1) * (b + 1) - (long long)r * 1000000007ll;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(&stg[slot][0]);
                if (mode == 0) {
                    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                                     reinterpret_cast<uint64_t>(&map)),
                                 "r"(x0), "r"(y), "r"(z), "r"(0), "r"(src)
                                 : "memory");
                } else if (y >= 0 && y < H) {
                    // 1-D bulk reduce: clip the row to the tensor by hand (start and size stay multiples of 16 bytes)
                    const int xa = max(x0, 0), xb = min(x0 + BOXW, W);
                    if (xb > xa) {
                        long long *dst = gbase + ((long long)z * H + y) * W + xa;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;" ::"l"(dst),
                                     "r"(src + 8u * (uint32_t)(xa - x0)), "r"((uint32_t)(xb - xa) * 8u)
                                     : "memory");
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0 && warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int inrange = argc > 2 ? atoi(argv[2]) : 0;
    const int W = 100, H = 13, D = 7, NB = 600, REPS = 20;
    std::vector<long long> h((size_t)W * H * D, 0), want((size_t)W * H * D, 0);
    for (size_t i = 0; i < h.size(); ++i) h[i] = want[i] = (long long)i * 3 - 1000;
    long long *d;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    CUtensorMap map;
    cuuint64_t dims[4] = {W, H, D, 1};
    cuuint64_t strides[3] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8, (cuuint64_t)W * H * D * 8};
    cuuint32_t box[4] = {BOXW, 1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_INT64, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    k<<<NB, 64>>>(map, d, mode, W, H, D, REPS);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    for (int b = 0; b < NB; ++b) {
        const int x0 = 12 * (b % 9) - 6, y = (b * 5) % (H + 2) - 1, z = (b * 3) % D;
        if (y < 0 || y >= H) continue;
        for (int rr = 0; rr < REPS; ++rr)
            for (int l = 0; l < BOXW; ++l) {
                const int x = x0 + l;
                if (x < 0 || x >= W) continue;
                want[((size_t)z * H + y) * W + x] += (long long)(x0 + l + 1) * (b + 1) - (long long)rr * 1000000007ll;
            }
    }
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (size_t i = 0; i < h.size(); ++i) bad += h[i] != want[i];
    printf("mismatches %ld of %zu\n", bad, h.size());
    // timing: how long do 1587 row-reduces per CTA take when every SM does it?
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148 * 4, 64>>>(map, d, mode, W, H, D, 69 * 16);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("148*4 CTAs x 1104 row reduces (24 int64 each): %.3f ms -> %.1f ns per reduce per SM-resident CTA\n", ms, ms * 1e6 / (4 * 1104));
    return 0;
}
