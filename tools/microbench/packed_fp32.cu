#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 add2(u64 a,u64 b){u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a,u64 b){u64 r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
template<int MODE>
__global__ void k(float* out, int iters, float s){
  float x[16];
  #pragma unroll
  for(int i=0;i<16;i++) x[i]=threadIdx.x*0.001f+i;
  if(MODE==0){ // scalar add+mul alternating (no fma)
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int i=0;i<16;i++) x[i]=__fadd_rn(x[i],s);
      #pragma unroll
      for(int i=0;i<16;i++) x[i]=__fmul_rn(x[i],s);
    }
  } else if (MODE==1){
    u64 p[8]; u64 ss=pk(s,s);
    #pragma unroll
    for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int i=0;i<8;i++) p[i]=add2(p[i],ss);
      #pragma unroll
      for(int i=0;i<8;i++) p[i]=mul2(p[i],ss);
    }
    #pragma unroll
    for(int i=0;i<8;i++) upk(p[i],x[2*i],x[2*i+1]);
  } else if (MODE==2){ // scalar ffma
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int r=0;r<2;r++)
      #pragma unroll
      for(int i=0;i<16;i++) x[i]=__fmaf_rn(x[i],s,s);
    }
  } else {
    u64 p[8]; u64 ss=pk(s,s);
    #pragma unroll
    for(int i=0;i<8;i++) p[i]=pk(x[2*i],x[2*i+1]);
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int r=0;r<2;r++)
      #pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],ss,ss);
    }
    #pragma unroll
    for(int i=0;i<8;i++) upk(p[i],x[2*i],x[2*i+1]);
  }
  float acc=0; 
  #pragma unroll
  for(int i=0;i<16;i++) acc+=x[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template<int MODE> void run(const char* name, int threads){
  float* out; cudaMalloc(&out, 148*8*1024*4);
  int iters=20000;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148,threads>>>(out,100,1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<148,threads>>>(out,iters,1.0001f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double laneops = 148.0*threads*iters*32.0;
  printf("%s threads=%d: %.3f ms, %.2f T lane-ops/s (peak 148*128*1.965e9=37.2)\n", name, threads, ms, laneops/ms/1e9);
  cudaFree(out);
}
int main(){
  for (int t : {128, 256, 448, 512, 1024}) {
    run<0>("scalar add/mul", t); run<1>("packed add2/mul2", t); run<2>("scalar ffma", t); run<3>("packed ffma2", t);
  }
  return 0;
}
