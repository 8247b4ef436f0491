// Stand-alone probes of cp.async.bulk.tensor box loads (not product code): which ranks / box widths / coordinates work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_check tma_check.cu;   ./tma_check RANK BOXW X0 [ELEMBYTES]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the TMA unit (async proxy) must see the initialised barrier
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, uint32_t bytes, unsigned char *out) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_raw);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar, bytes);
        const uint64_t m = reinterpret_cast<uint64_t>(&tmap);
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(mbar), "r"(c0), "r"(c1) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(m), "r"(mbar), "r"(c0), "r"(c1), "r"(c2) : "memory");
        if (RANK == 4)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(m), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    mbar_wait(mbar, 0);
    for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = s_raw[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int rank = argc > 1 ? atoi(argv[1]) : 4, boxw = argc > 2 ? atoi(argv[2]) : 24, x0 = argc > 3 ? atoi(argv[3]) : 5;
    const int eb = argc > 4 ? atoi(argv[4]) : 2;
    const int E = argc > 5 ? atoi(argv[5]) : 8;
    const int D = 40, H = 48, W = 64, NV = 2;
    const size_t n = (size_t)NV * D * H * W;
    std::vector<unsigned char> h(n * eb);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 2654435761u >> 13);
    unsigned char *d, *dout;
    CK(cudaMalloc(&d, h.size())); CK(cudaMalloc(&dout, 1 << 16));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    void *ptr = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)ptr;
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)NV};
    cuuint64_t strides[3] = {(cuuint64_t)W * eb, (cuuint64_t)W * H * eb, (cuuint64_t)W * H * D * eb};
    cuuint32_t box[4] = {(cuuint32_t)boxw, (cuuint32_t)E, (cuuint32_t)E, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (rank == 2) { dims[1] = (cuuint64_t)H * D * NV; }
    if (rank == 3) { dims[2] = (cuuint64_t)D * NV; }
    CUresult r = fn(&map, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d boxw %d x0 %d elem %d B: encode %d; ", rank, boxw, x0, eb, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    uint32_t bytes = (uint32_t)boxw * E * eb * (rank >= 3 ? E : 1);
    cudaError_t e;
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    if (rank == 2) { k<2><<<1, 128, 64 * 1024>>>(map, x0, 3, 0, 0, bytes, dout); }
    if (rank == 3) { k<3><<<1, 128, 64 * 1024>>>(map, x0, 3, 2, 0, bytes, dout); }
    if (rank == 4) { k<4><<<1, 128, 64 * 1024>>>(map, x0, 3, 2, 1, bytes, dout); }
    e = cudaDeviceSynchronize();
    printf("run: %s; ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    std::vector<unsigned char> o(bytes);
    CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
    long bad = 0;
    const int nz = rank >= 3 ? E : 1;
    for (int z = 0; z < nz; ++z) for (int y = 0; y < E; ++y) for (int x = 0; x < boxw; ++x) for (int b = 0; b < eb; ++b) {
        const long gx = x0 + x, gy = 3 + y, gz = (rank >= 3 ? 2 + z : 0), gv = rank == 4 ? 1 : 0;
        bool in = gx >= 0 && gx < W;
        size_t idx;
        if (rank == 2) { in = in && gy < (long)H * D * NV; idx = ((size_t)gy * W + gx); }
        else if (rank == 3) { in = in && gy < H && gz < (long)D * NV; idx = (((size_t)gz * H + gy) * W + gx); }
        else { in = in && gy < H && gz < D; idx = ((((size_t)gv * D + gz) * H + gy) * W + gx); }
        const unsigned char want = in ? h[idx * eb + b] : 0;
        bad += o[(((size_t)z * E + y) * boxw + x) * eb + b] != want;
    }
    printf("mismatches %ld of %u\n", bad, bytes);
    return 0;
}
