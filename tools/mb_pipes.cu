// Issue-rate microbenchmarks that decide the matcher design (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_pipes mb_pipes.cu
// Each variant: 8 independent chains per thread, 1024 threads/CTA, 2 CTAs/SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

template <int W>
__global__ void __launch_bounds__(1024) k(int iters, unsigned* sink) {
  unsigned a[8]; int s[8]; float f[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 2654435761u + i; s[i] = (int)a[i]; f[i] = (float)(threadIdx.x + i) * 1e-3f; }
  const unsigned m = threadIdx.x | 1u, c = blockIdx.x + 3u;
  const int ms = (int)(threadIdx.x * 0x01010101u) ^ 0x7f3f1f0f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (W == 0) a[i] = a[i] * m + c;                                   // IMAD
        if (W == 1) s[i] = __dp4a(ms, (int)(c + i), s[i]);                 // IDP.4A s8*s8 (b varies little)
        if (W == 2) a[i] = __dp4a((unsigned)ms, c + i, a[i]);              // IDP.4A u8*u8
        if (W == 3) s[i] = __dp2a_lo(ms, (int)(c + i), s[i]);              // IDP.2A
        if (W == 4) { unsigned p = __byte_perm(a[(i + 1) & 7] , c, 0x4321); s[i] = __dp4a((int)p, ms, s[i]); }  // PRMT + IDP.4A
        if (W == 5) a[i] = __funnelshift_r(a[i], c, 8) ^ m;                // SHF + LOP (ALU only)
        if (W == 6) f[i] = __fmaf_rn(f[i], 1.0000001f, 1e-7f);             // FFMA
        if (W == 7) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 3)); // SHFL
        if (W == 8) { long long q = __float2ll_rn(f[i]); f[i] += (float)(unsigned)q * 1e-9f + 1.0f; }  // F2I.S64 (+I2F)
        if (W == 9) { int q = __float2int_rn(f[i]); f[i] = __int_as_float((q & 0x7fffff) | 0x3f800000); }  // F2I.S32 + LOP
      }
    }
  }
  unsigned r = 0;
  for (int i = 0; i < 8; ++i) r ^= a[i] ^ (unsigned)s[i] ^ __float_as_uint(f[i]);
  if (r == 0x12345678u) sink[0] = r;
}

template <int W> void run(const char* name, double per_op, unsigned* sink, int sms) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int blocks = sms * 2, iters = 2048;
  k<W><<<blocks, 1024>>>(64, sink);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); k<W><<<blocks, 1024>>>(iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * 1024.0 * iters * 16.0 * 8.0;
    double rate = ops / (ms * 1e-3); if (rate > best) best = rate;
  }
  printf("%-28s %8.2f Tlane-op/s  = %6.1f lanes/clk/SM   (%s)\n", name, best / 1e12, best / (sms * 1.965e9), per_op > 1 ? "x4 MACs per op" : "");
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  unsigned* sink; CK(cudaMalloc(&sink, 64));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  int n = p.multiProcessorCount;
  run<0>("IMAD", 1, sink, n); run<1>("IDP.4A s8", 4, sink, n); run<2>("IDP.4A u8", 4, sink, n); run<3>("IDP.2A", 2, sink, n);
  run<4>("PRMT + IDP.4A (pair)", 4, sink, n); run<5>("SHF + LOP (pair)", 1, sink, n); run<6>("FFMA", 1, sink, n);
  run<7>("SHFL.BFLY", 1, sink, n); run<8>("F2I.S64 + I2F + FADD..", 1, sink, n); run<9>("F2I.S32 + LOP", 1, sink, n);
  printf("done\n"); return 0;
}
