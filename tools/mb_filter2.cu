// Microbenchmarks behind the round-2 filter redesign (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_filter2 mb_filter2.cu
// 1. packed fp32 (FFMA2 / FADD2) issue rate against scalar FFMA
// 2. red.shared.add.u32 cost by address pattern (4x4x2 block pattern vs 32 consecutive words),
//    against the plain LDS + IADD + STS sequence on the same addresses
// 3. shared-memory transposition voxel-per-lane -> block-per-lane (64 STS.32 + 16 LDS.128, row stride 68 words)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a,u64 b,u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a,u64 b){ u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r; }

template<int MODE> // 0 FFMA  1 FFMA2  2 FADD2  3 FFMA2 + FFMA interleaved
__global__ void k_fp(int iters, float* sink){
  float a[8]; u64 p[8];
  for(int i=0;i<8;++i){ a[i]=threadIdx.x*0.001f+i; float2 t=make_float2(a[i],a[i]+1.f); p[i]=*reinterpret_cast<u64*>(&t); }
  const float c=1.0001f; float2 cc=make_float2(c,c); const u64 c2=*reinterpret_cast<u64*>(&cc);
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int u=0;u<8;++u){
#pragma unroll
      for(int i=0;i<8;++i){
        if(MODE==0) a[i]=__fmaf_rn(a[i],c,c);
        if(MODE==1) p[i]=ffma2(p[i],c2,c2);
        if(MODE==2) p[i]=fadd2(p[i],c2);
        if(MODE==3){ p[i]=ffma2(p[i],c2,c2); a[i]=__fmaf_rn(a[i],c,c); }
      }
    }
  }
  float s=0; for(int i=0;i<8;++i){ float2 t=*reinterpret_cast<float2*>(&p[i]); s+=a[i]+t.x+t.y; }
  if(s==1.2345f) sink[0]=s;
}

constexpr int SY=24, SZ=556, PLANES=17, TILE=PLANES*SZ;
// MODE 0: RED block pattern (lane = zh,y,x; two registers = two planes)   1: RED 32 consecutive words
//      2: LDS+IADD+STS block pattern   3: LDS+IADD+STS consecutive   4: RED block pattern, half of the lanes predicated off
template<int MODE>
__global__ void k_red(int iters, unsigned* sink){
  extern __shared__ uint32_t sm[];
  for(int i=threadIdx.x;i<TILE;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const int lx=lane&3, ly=(lane>>2)&3, zh=lane>>4;
  const uint32_t base=(uint32_t)__cvta_generic_to_shared(sm);
  uint32_t off[16];
  for(int i=0;i<16;++i){
    uint32_t h=hash32(blockIdx.x*977u+warp*131u+i*17u+1u);
    const int oz=h%11, oy=(h>>8)%20, ox=(h>>16)%20;
    if(MODE==1||MODE==3) off[i]=4u*(uint32_t)(oz*SZ+oy*SY+ (ox%16) + lane);           // one plane row run (may wrap rows: still 32 consecutive words)
    else off[i]=4u*(uint32_t)((oz+2*zh)*SZ+(oy+ly)*SY+ox+lx);
  }
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int i=0;i<16;++i){
      const uint32_t a0=base+off[i], a1=a0+4u*SZ;
      if(MODE==0||MODE==1){
        asm volatile("red.shared.add.u32 [%0], %1;"::"r"(a0),"r"(3u):"memory");
        asm volatile("red.shared.add.u32 [%0], %1;"::"r"(a1),"r"(5u):"memory");
      } else if(MODE==4){
        if(lane&1){
        asm volatile("red.shared.add.u32 [%0], %1;"::"r"(a0),"r"(3u):"memory");
        asm volatile("red.shared.add.u32 [%0], %1;"::"r"(a1),"r"(5u):"memory");
        }
      } else {
        uint32_t v0,v1;
        asm volatile("ld.shared.u32 %0, [%1];":"=r"(v0):"r"(a0):"memory");
        asm volatile("ld.shared.u32 %0, [%1];":"=r"(v1):"r"(a1):"memory");
        v0+=3u; v1+=5u;
        asm volatile("st.shared.u32 [%0], %1;"::"r"(a0),"r"(v0):"memory");
        asm volatile("st.shared.u32 [%0], %1;"::"r"(a1),"r"(v1):"memory");
      }
    }
  }
  __syncthreads();
  unsigned s=0; for(int i=threadIdx.x;i<TILE;i+=blockDim.x) s+=sm[i];
  if(s==12345u) sink[0]=s;
}

// transposition: per warp a buffer of 32 rows x 68 words.  write: 64 STS.32 (row j = member, 32 consecutive voxels per
// instruction), read: lane j takes its row as 16 LDS.128.  Then the way back: 16 STS.128 + 64 LDS.32.
__global__ void k_tr(int iters, float* sink){
  extern __shared__ float4 smf4[];
  float* sm=reinterpret_cast<float*>(smf4);
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  float* T=sm+warp*(32*68);
  float v[64];
  for(int i=0;i<64;++i) v[i]=lane+i*0.5f;
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int j=0;j<32;++j){ T[j*68+lane]=v[2*j]; T[j*68+32+lane]=v[2*j+1]; }
    __syncwarp();
#pragma unroll
    for(int q=0;q<16;++q){ float4 t=*reinterpret_cast<float4*>(T+lane*68+4*q); v[4*q]=t.x+1.f; v[4*q+1]=t.y; v[4*q+2]=t.z; v[4*q+3]=t.w; }
    __syncwarp();
#pragma unroll
    for(int q=0;q<16;++q){ *reinterpret_cast<float4*>(T+lane*68+4*q)=make_float4(v[4*q],v[4*q+1],v[4*q+2],v[4*q+3]); }
    __syncwarp();
#pragma unroll
    for(int j=0;j<32;++j){ v[2*j]=T[j*68+lane]; v[2*j+1]=T[j*68+32+lane]+1.f; }
    __syncwarp();
  }
  float s=0; for(int i=0;i<64;++i) s+=v[i];
  if(s==1.2345f) sink[0]=s;
}

template<class F> float timeit(F f, int reps=3){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int i=0;i<reps;++i){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  const int nsm=pr.multiProcessorCount; const double clk=pr.clockRate*1e3;
  printf("device %s, %d SMs, clock %.0f MHz (cycle figures assume this clock)\n", pr.name, nsm, clk/1e6);
  float* sinkf; unsigned* sinku; CK(cudaMalloc(&sinkf,64)); CK(cudaMalloc(&sinku,64));
  {
    const int iters=20000, thr=512, cta=nsm*2;
    const char* nm[4]={"FFMA","FFMA2","FADD2","FFMA2+FFMA"};
    for(int m=0;m<4;++m){
      float ms=0;
      if(m==0) ms=timeit([&]{k_fp<0><<<cta,thr>>>(iters,sinkf);});
      if(m==1) ms=timeit([&]{k_fp<1><<<cta,thr>>>(iters,sinkf);});
      if(m==2) ms=timeit([&]{k_fp<2><<<cta,thr>>>(iters,sinkf);});
      if(m==3) ms=timeit([&]{k_fp<3><<<cta,thr>>>(iters,sinkf);});
      const double inst=(double)iters*64*(m==3?2:1)*thr*cta;   // thread-instructions
      printf("%-12s %7.3f ms  %6.1f thread-instr/clk/SM  (%s)\n", nm[m], ms, inst/(ms*1e-3)/clk/nsm, m==0?"1 fma each":m==3?"1.5 fma each on average":"2 flops-lanes each");
    }
  }
  {
    const size_t smem=(size_t)TILE*4;
    const char* nm[5]={"RED block pattern","RED 32 consecutive","LDS+IADD+STS block","LDS+IADD+STS consecutive","RED block, odd lanes only"};
    for(int warps: {8,16,32}){
      const int iters=4000;
      for(int m=0;m<5;++m){
        float ms=0;
        if(m==0){ CK(cudaFuncSetAttribute(k_red<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_red<0><<<nsm,warps*32,smem>>>(iters,sinku);}); }
        if(m==1){ CK(cudaFuncSetAttribute(k_red<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_red<1><<<nsm,warps*32,smem>>>(iters,sinku);}); }
        if(m==2){ CK(cudaFuncSetAttribute(k_red<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_red<2><<<nsm,warps*32,smem>>>(iters,sinku);}); }
        if(m==3){ CK(cudaFuncSetAttribute(k_red<3>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_red<3><<<nsm,warps*32,smem>>>(iters,sinku);}); }
        if(m==4){ CK(cudaFuncSetAttribute(k_red<4>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_red<4><<<nsm,warps*32,smem>>>(iters,sinku);}); }
        const double ops=(double)iters*32*warps;  // warp-level updates (one RED, or one LDS+IADD+STS) per SM
        printf("%-28s warps/SM %2d: %7.3f ms -> %5.2f clk per warp-wide update per SM\n", nm[m], warps, ms, ms*1e-3*clk/ops);
      }
    }
  }
  {
    for(int warps: {8,12,16}){
      const size_t smem=(size_t)warps*32*68*4; const int iters=2000;
      CK(cudaFuncSetAttribute(k_tr,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
      float ms=timeit([&]{k_tr<<<nsm,warps*32,smem>>>(iters,sinkf);});
      printf("transposition round trip (64 STS.32 + 16 LDS.128 + 16 STS.128 + 64 LDS.32) warps/SM %2d: %7.3f ms -> %6.1f clk per round trip per SM (256 wavefronts of 128 B)\n",
             warps, ms, ms*1e-3*clk/((double)iters*warps));
    }
  }
  printf("done\n");
  return 0;
}
