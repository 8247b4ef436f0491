// Microbenchmarks that decide the aggregation design (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_atomics mb_atomics.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// ---- 1. shared-memory atomics / RMW, conflict-free block pattern -----------------
// tile 23 x 24(pad) x planes stride 580; "block" = 4x4x4 voxels at a random origin; lane = (z:2, yhi:1, x:2), 2 regs (ylo)
constexpr int SY=24, SZ=580, TILE=23*SZ;
template<int MODE>  // 0: ATOMS.ADD u32   1: float atomicAdd (CAS loop)   2: non-atomic LDS/FADD/STS   3: u64 atomicAdd
__global__ void k_smem(int iters, unsigned* sink){
  extern __shared__ uint32_t sm[];
  uint32_t* a = sm; uint32_t* b = sm + TILE;
  for(int i=threadIdx.x;i<2*TILE;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
  const int lz=lane>>3, lyh=(lane>>2)&1, lx=lane&3;
  const int off0 = lz*SZ + (2*lyh)*SY + lx, off1 = off0+SY;
  uint32_t rng = hash32(blockIdx.x*977u+warp*131u+1u);
  for(int it=0; it<iters; ++it){
    rng = hash32(rng+it);
    const int oz = rng%20, oy=(rng>>8)%20, ox=(rng>>16)%20;
    const int base = oz*SZ+oy*SY+ox;
    if(MODE==0){
      atomicAdd(&a[base+off0], 3u); atomicAdd(&a[base+off1], 5u);
      atomicAdd(&b[base+off0], 7u); atomicAdd(&b[base+off1], 9u);
    } else if(MODE==1){
      atomicAdd((float*)&a[base+off0], 3.f); atomicAdd((float*)&a[base+off1], 5.f);
      atomicAdd((float*)&b[base+off0], 7.f); atomicAdd((float*)&b[base+off1], 9.f);
    } else if(MODE==2){
      float* fa=(float*)a; float* fb=(float*)b;
      float v0=fa[base+off0], v1=fa[base+off1], v2=fb[base+off0], v3=fb[base+off1];
      fa[base+off0]=v0+3.f; fa[base+off1]=v1+5.f; fb[base+off0]=v2+7.f; fb[base+off1]=v3+9.f;
      __syncwarp();
    } else {
      unsigned long long* la=(unsigned long long*)sm;  // [TILE] u64
      atomicAdd(&la[base+off0], 3ull); atomicAdd(&la[base+off1], 5ull);
    }
  }
  __syncthreads();
  unsigned s=0; for(int i=threadIdx.x;i<2*TILE;i+=blockDim.x) s+=sm[i];
  if(s==12345u) sink[0]=s;
}

// ---- 2. global reductions, BM4D-like locality ---------------------------------------
// Volume S^3 of float2 (num,den) (or 2x int64).  Each warp walks reference blocks on the step-3 grid;
// for each ref it adds K blocks at random origins inside the 11^3 window.
// PAT 0: lane = block (lanes<K), 64 red.v2.f32 per lane (current kernel)
// PAT 1: lane = voxel pair: for k: lane (z:2,y:2,xh:1) issues one red.v4.f32 (two voxels, 16 B)   [needs 16B alignment -> x origin even only; test only]
// PAT 2: lane = voxel: for k: 2 x red.v2.f32, lanes (z:2,yhi:1,x:2) -> 4 lanes share a 32B row
// PAT 3: lane = block, 2 x 64-bit integer atomics per voxel (current deterministic mode)
// PAT 4: lane = voxel, 2x int64 atomics, 4 lanes contiguous
template<int PAT>
__global__ void k_glob(float2* acc, long long* numq, long long* denq, int S, int K, long long nrefs, int nrx){
  const int lane=threadIdx.x&31;
  const long long wid = (long long)blockIdx.x*(blockDim.x>>5)+(threadIdx.x>>5);
  const long long nw = (long long)gridDim.x*(blockDim.x>>5);
  for(long long r=wid; r<nrefs; r+=nw){
    const int ix=(int)(r%nrx), iy=(int)((r/nrx)%nrx), iz=(int)(r/((long long)nrx*nrx));
    const int oz=iz*3, oy=iy*3, ox=ix*3;
    if(PAT==0||PAT==3){
      if(lane<K){
        uint32_t h=hash32((uint32_t)r*33u+lane);
        int cz=min(max(oz-5+(int)(h%11),0),S-4), cy=min(max(oy-5+(int)((h>>8)%11),0),S-4), cx=min(max(ox-5+(int)((h>>16)%11),0),S-4);
        for(int z=0;z<4;++z)for(int y=0;y<4;++y)for(int x=0;x<4;++x){
          long long a=((long long)(cz+z)*S+(cy+y))*S+cx+x;
          if(PAT==0) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};"::"l"(acc+a),"f"(1.5f),"f"(0.25f):"memory");
          else { atomicAdd((unsigned long long*)(numq+a), 12345ull); atomicAdd((unsigned long long*)(denq+a), 777ull); }
        }
      }
    } else {
      for(int k=0;k<K;++k){
        uint32_t h=hash32((uint32_t)r*33u+k);
        int cz=min(max(oz-5+(int)(h%11),0),S-4), cy=min(max(oy-5+(int)((h>>8)%11),0),S-4), cx=min(max(ox-5+(int)((h>>16)%11),0),S-4);
        if(PAT==1){
          cx &= ~1;
          const int z=lane>>3, y=(lane>>1)&3, xh=lane&1;
          long long a=((long long)(cz+z)*S+(cy+y))*S+cx+2*xh;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"::"l"(acc+a),"f"(1.5f),"f"(0.25f),"f"(2.5f),"f"(0.125f):"memory");
        } else {
          const int z=lane>>3, yh=(lane>>2)&1, x=lane&3;
          long long a0=((long long)(cz+z)*S+(cy+2*yh))*S+cx+x, a1=a0+S;
          if(PAT==2){
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};"::"l"(acc+a0),"f"(1.5f),"f"(0.25f):"memory");
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};"::"l"(acc+a1),"f"(1.5f),"f"(0.25f):"memory");
          } else {
            atomicAdd((unsigned long long*)(numq+a0), 12345ull); atomicAdd((unsigned long long*)(denq+a0), 777ull);
            atomicAdd((unsigned long long*)(numq+a1), 12345ull); atomicAdd((unsigned long long*)(denq+a1), 777ull);
          }
        }
      }
    }
  }
}

// gather: lane = block, 64 scalar __ldg per lane (current) vs lane=voxel
template<int PAT>
__global__ void k_gather(const float* src, float* out, int S, int K, long long nrefs, int nrx){
  const int lane=threadIdx.x&31;
  const long long wid = (long long)blockIdx.x*(blockDim.x>>5)+(threadIdx.x>>5);
  const long long nw = (long long)gridDim.x*(blockDim.x>>5);
  float s=0.f;
  for(long long r=wid; r<nrefs; r+=nw){
    const int ix=(int)(r%nrx), iy=(int)((r/nrx)%nrx), iz=(int)(r/((long long)nrx*nrx));
    const int oz=iz*3, oy=iy*3, ox=ix*3;
    if(PAT==0){
      if(lane<K){
        uint32_t h=hash32((uint32_t)r*33u+lane);
        int cz=min(max(oz-5+(int)(h%11),0),S-4), cy=min(max(oy-5+(int)((h>>8)%11),0),S-4), cx=min(max(ox-5+(int)((h>>16)%11),0),S-4);
        #pragma unroll
        for(int z=0;z<4;++z)
        #pragma unroll
        for(int y=0;y<4;++y)
        #pragma unroll
        for(int x=0;x<4;++x) s+=__ldg(src+((long long)(cz+z)*S+(cy+y))*S+cx+x);
      }
    } else {
      #pragma unroll 4
      for(int k=0;k<K;++k){
        uint32_t h=hash32((uint32_t)r*33u+k);
        int cz=min(max(oz-5+(int)(h%11),0),S-4), cy=min(max(oy-5+(int)((h>>8)%11),0),S-4), cx=min(max(ox-5+(int)((h>>16)%11),0),S-4);
        const int z=lane>>3, yh=(lane>>2)&1, x=lane&3;
        long long a0=((long long)(cz+z)*S+(cy+2*yh))*S+cx+x;
        s+=__ldg(src+a0)+__ldg(src+a0+S);
      }
    }
  }
  if(s==1.2345f) out[0]=s;
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
  printf("device %s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
  const int nsm=pr.multiProcessorCount;
  unsigned* sink; CK(cudaMalloc(&sink,64));
  // ---- smem
  const size_t smem=2*TILE*4;
  const char* names[4]={"ATOMS.ADD.u32","atomicAdd float (CAS)","LDS/FADD/STS non-atomic","atomicAdd u64 (CAS)"};
  for(int warps: {4,8,16,32}){
    const int iters=20000;
    auto run=[&](int mode){
      float ms=0;
      if(mode==0){ CK(cudaFuncSetAttribute(k_smem<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_smem<0><<<nsm,warps*32,smem>>>(iters,sink);}); }
      if(mode==1){ CK(cudaFuncSetAttribute(k_smem<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_smem<1><<<nsm,warps*32,smem>>>(iters,sink);}); }
      if(mode==2){ CK(cudaFuncSetAttribute(k_smem<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_smem<2><<<nsm,warps*32,smem>>>(iters,sink);}); }
      if(mode==3){ CK(cudaFuncSetAttribute(k_smem<3>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{k_smem<3><<<nsm,warps*32,smem>>>(iters,sink);}); }
      const double vox = (mode==3?1.0:1.0)*64.0*iters*warps;  // block-voxel updates (num+den both, mode 3: one u64 array only)
      // assume 1.9 GHz for cycles estimate
      printf("smem %-26s warps/SM %2d: %8.3f ms  -> %.2f clk per 64-voxel block update per SM (@1.9GHz), %.1f Gvoxel-updates/s chip\n", names[mode], warps, ms, ms*1e-3*1.9e9/(iters*warps), vox*nsm/(ms*1e-3)/1e9);
    };
    for(int m=0;m<4;++m) run(m);
  }
  // ---- global
  const int S=768; const long long V=(long long)S*S*S; const int nrx=(S-4)/3+1; const long long nrefs=(long long)nrx*nrx*nrx;
  float2* acc; long long *nq,*dq; float* src;
  CK(cudaMalloc(&acc,V*8)); CK(cudaMalloc(&nq,V*8)); CK(cudaMalloc(&dq,V*8)); CK(cudaMalloc(&src,V*4));
  CK(cudaMemset(acc,0,V*8)); CK(cudaMemset(nq,0,V*8)); CK(cudaMemset(dq,0,V*8)); CK(cudaMemset(src,0,V*4));
  const char* pn[5]={"lane=block red.v2.f32 (current fast)","lane=voxelpair red.v4.f32","lane=voxel 4-lane rows red.v2.f32","lane=block 2x int64 atomics (current det)","lane=voxel 2x int64 atomics"};
  for(int K: {16,32}){
    for(int pat=0;pat<5;++pat){
      const int blocks=nsm*16, thr=128;
      float ms=0;
      if(pat==0) ms=timeit([&]{k_glob<0><<<blocks,thr>>>(acc,nq,dq,S,K,nrefs,nrx);});
      if(pat==1) ms=timeit([&]{k_glob<1><<<blocks,thr>>>(acc,nq,dq,S,K,nrefs,nrx);});
      if(pat==2) ms=timeit([&]{k_glob<2><<<blocks,thr>>>(acc,nq,dq,S,K,nrefs,nrx);});
      if(pat==3) ms=timeit([&]{k_glob<3><<<blocks,thr>>>(acc,nq,dq,S,K,nrefs,nrx);});
      if(pat==4) ms=timeit([&]{k_glob<4><<<blocks,thr>>>(acc,nq,dq,S,K,nrefs,nrx);});
      printf("glob K=%d %-42s: %8.3f ms for %lld refs -> %.2f ns/ref, %.1f Gvoxel-updates/s; scaled to 1024^3 (39.65M refs): %.1f ms\n", K, pn[pat], ms, nrefs, ms*1e6/nrefs, (double)nrefs*K*64/(ms*1e-3)/1e9, ms*39651821.0/nrefs);
    }
    for(int pat=0;pat<2;++pat){
      const int blocks=nsm*16, thr=128; float ms;
      if(pat==0) ms=timeit([&]{k_gather<0><<<blocks,thr>>>(src,(float*)sink,S,K,nrefs,nrx);});
      else ms=timeit([&]{k_gather<1><<<blocks,thr>>>(src,(float*)sink,S,K,nrefs,nrx);});
      printf("gather K=%d %s: %8.3f ms -> scaled to 1024^3: %.1f ms\n", K, pat==0?"lane=block scalar ldg":"lane=voxel 4-lane rows", ms, ms*39651821.0/nrefs);
    }
  }
  printf("done\n");
  return 0;
}
