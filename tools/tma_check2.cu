// TMA probe with the official libcu++ wrappers (CUDA programming guide example) — is the descriptor good?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdint.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 32, BH = 8;
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, uint16_t *out, int mode) {
    __shared__ alignas(128) uint16_t smem[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem, &tmap, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem[i / BW][i % BW];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int H = 256, W = 64;
    std::vector<uint16_t> h((size_t)H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i + 1);
    uint16_t *d, *dout;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&dout, BW * BH * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    CUtensorMap map;
    cuuint64_t dims[2] = {W, H};
    cuuint64_t strides[1] = {W * 2};
    cuuint32_t box[2] = {BW, BH};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    k<<<1, 128>>>(map, 8, 3, dout, 0);
    printf("run: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<uint16_t> o(BW * BH);
    cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) bad += o[y * BW + x] != h[(size_t)(3 + y) * W + 8 + x];
    printf("mismatches %ld of %d; first values got %u %u %u want %u %u %u\n", bad, BW * BH, o[0], o[1], o[2], h[3 * W + 8], h[3 * W + 9], h[3 * W + 10]);
    return 0;
}
