// Checks on the GPU box that the packed fp32 instructions round like the scalar IEEE forms and that the
// correction sequence used for W = a / (a + s2) gives the correctly rounded quotient (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o packed_check packed_check.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up(u64 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float rcp_approx(float d) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d)); return r; }

// out[0..3]: mismatch counts of add2, mul2, fma2, division; out[4..]: first failing division operands
__global__ void k(const float *a, const float *b, const float *c, int n, float s2, unsigned *out, float *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = a[i], y = b[i], z = c[i];
    float lo, hi;
    up(add2(pk(x, y), pk(z, x)), lo, hi);
    if (__float_as_uint(lo) != __float_as_uint(__fadd_rn(x, z)) || __float_as_uint(hi) != __float_as_uint(__fadd_rn(y, x))) atomicAdd(out + 0, 1u);
    up(mul2(pk(x, y), pk(z, x)), lo, hi);
    if (__float_as_uint(lo) != __float_as_uint(__fmul_rn(x, z)) || __float_as_uint(hi) != __float_as_uint(__fmul_rn(y, x))) atomicAdd(out + 1, 1u);
    up(fma2(pk(x, y), pk(z, x), pk(y, z)), lo, hi);
    if (__float_as_uint(lo) != __float_as_uint(__fmaf_rn(x, z, y)) || __float_as_uint(hi) != __float_as_uint(__fmaf_rn(y, x, z))) atomicAdd(out + 2, 1u);
    // the Wiener attenuation on y2 = x^2 (x spans many magnitudes)
    const u64 yn = pk(x, y);
    const u64 y2 = mul2(yn, yn);
    const u64 nd = sub2(pk(-s2, -s2), y2);
    float d0, d1;
    up(nd, d0, d1);
    const u64 r0 = pk(rcp_approx(-d0), rcp_approx(-d1));
    const u64 e0 = fma2(nd, r0, pk(1.0f, 1.0f));
    const u64 r1 = fma2(r0, e0, r0);
    const u64 q0 = mul2(y2, r1);
    const u64 e1 = fma2(nd, q0, y2);
    const u64 ww = fma2(r1, e1, q0);
    up(ww, lo, hi);
    const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y);
    const float t0 = __fdiv_rn(xx, __fadd_rn(xx, s2)), t1 = __fdiv_rn(yy, __fadd_rn(yy, s2));
    if (__float_as_uint(lo) != __float_as_uint(t0) || __float_as_uint(hi) != __float_as_uint(t1)) {
        if (atomicAdd(out + 3, 1u) == 0) {
            bad[0] = x; bad[1] = lo; bad[2] = t0; bad[3] = y; bad[4] = hi; bad[5] = t1;
        }
    }
}

int main() {
    const int n = 1 << 24;
    std::vector<float> a(n), b(n), c(n);
    srand(1);
    for (int i = 0; i < n; ++i) {
        auto rnd = [&]() {
            const double m = (rand() / (double)RAND_MAX) * 2.0 - 1.0;
            const int e = rand() % 60 - 40;
            return (float)std::ldexp(m, e);
        };
        a[i] = rnd(); b[i] = rnd(); c[i] = rnd();
    }
    float *da, *db, *dc, *dbad; unsigned *dout;
    cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dc, n * 4); cudaMalloc(&dout, 64); cudaMalloc(&dbad, 64);
    cudaMemcpy(da, a.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dc, c.data(), n * 4, cudaMemcpyHostToDevice);
    for (float s2 : {576.0f, 100.0f, 0.0025f, 1.0f}) {
        cudaMemset(dout, 0, 64); cudaMemset(dbad, 0, 64);
        k<<<n / 256, 256>>>(da, db, dc, n, s2, dout, dbad);
        unsigned out[4]; float bad[6];
        cudaMemcpy(out, dout, 16, cudaMemcpyDeviceToHost); cudaMemcpy(bad, dbad, 24, cudaMemcpyDeviceToHost);
        printf("s2 %g: mismatches add2 %u mul2 %u fma2 %u division %u of %d", s2, out[0], out[1], out[2], out[3], n);
        if (out[3]) printf("  first: x %.9g got %.9g want %.9g | y %.9g got %.9g want %.9g", bad[0], bad[1], bad[2], bad[3], bad[4], bad[5]);
        printf("\n");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
