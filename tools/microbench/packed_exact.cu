#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 add2(u64 a,u64 b){u64 r; asm("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 sub2(u64 a,u64 b){u64 r; asm("sub.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 mul2(u64 a,u64 b){u64 r; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
// out[0..]: mismatch counters per test
__global__ void k(const float* x, int n, unsigned* bad){
  int i = blockIdx.x*blockDim.x+threadIdx.x;
  if (4*i+7 >= n) return;
  float a=x[4*i], b=x[4*i+1], c=x[4*i+2], d=x[4*i+3], e=x[4*i+4], f=x[4*i+5];
  u64 A=pk(a,b), B=pk(c,d), C=pk(e,f);
  float lo,hi;
  upk(add2(A,B),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fadd_rn(a,c))||__float_as_uint(hi)!=__float_as_uint(__fadd_rn(b,d))) atomicAdd(bad+0,1);
  upk(sub2(A,B),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fsub_rn(a,c))||__float_as_uint(hi)!=__float_as_uint(__fsub_rn(b,d))) atomicAdd(bad+1,1);
  upk(mul2(A,B),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fmul_rn(a,c))||__float_as_uint(hi)!=__float_as_uint(__fmul_rn(b,d))) atomicAdd(bad+2,1);
  upk(fma2(A,B,C),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fmaf_rn(a,c,e))||__float_as_uint(hi)!=__float_as_uint(__fmaf_rn(b,d,f))) atomicAdd(bad+3,1);
  // chained mul then add (must NOT be fused)
  upk(add2(mul2(A,B),C),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fadd_rn(__fmul_rn(a,c),e))||__float_as_uint(hi)!=__float_as_uint(__fadd_rn(__fmul_rn(b,d),f))) atomicAdd(bad+4,1);
  upk(sub2(mul2(A,B),C),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fsub_rn(__fmul_rn(a,c),e))||__float_as_uint(hi)!=__float_as_uint(__fsub_rn(__fmul_rn(b,d),f))) atomicAdd(bad+5,1);
  upk(sub2(C,mul2(A,B)),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fsub_rn(e,__fmul_rn(a,c)))||__float_as_uint(hi)!=__float_as_uint(__fsub_rn(f,__fmul_rn(b,d)))) atomicAdd(bad+6,1);
  // div3
  { const float r=0.3333333432674407958984375f; u64 R=pk(r,r); u64 q=mul2(A,R); u64 ee=fma2(pk(-3.f,-3.f),q,A); u64 qq=fma2(ee,R,q); upk(qq,lo,hi);
    if (__float_as_uint(lo)!=__float_as_uint(__fdiv_rn(a,3.f))||__float_as_uint(hi)!=__float_as_uint(__fdiv_rn(b,3.f))) atomicAdd(bad+7,1); }
  // (a+c)-a pattern
  upk(sub2(add2(A,B),A),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fsub_rn(__fadd_rn(a,c),a))||__float_as_uint(hi)!=__float_as_uint(__fsub_rn(__fadd_rn(b,d),b))) atomicAdd(bad+8,1);
  // mul by 4 then fma
  upk(fma2(pk(4.f,4.f),B,mul2(pk(10.f,10.f),A)),lo,hi); if (__float_as_uint(lo)!=__float_as_uint(__fadd_rn(__fmul_rn(4.f,c),__fmul_rn(10.f,a)))||__float_as_uint(hi)!=__float_as_uint(__fadd_rn(__fmul_rn(4.f,d),__fmul_rn(10.f,b)))) atomicAdd(bad+9,1);
}
int main(){
  int n = 1<<24; float* h=(float*)malloc(n*4);
  srand(1);
  for(int i=0;i<n;i++){ int m=rand()%4; float v=(float)rand()/RAND_MAX; 
    if(m==0) h[i]=250.f+v*60.f; else if(m==1) h[i]=(v-0.5f)*2e-3f; else if (m==2) h[i]=v*1e-38f; else h[i]=(v-0.5f)*1e3f; }
  float* d; unsigned* bad; cudaMalloc(&d,n*4); cudaMalloc(&bad,64); cudaMemset(bad,0,64);
  cudaMemcpy(d,h,n*4,cudaMemcpyHostToDevice);
  k<<<(n/4+255)/256,256>>>(d,n,bad);
  unsigned hb[16]; cudaMemcpy(hb,bad,64,cudaMemcpyDeviceToHost);
  const char* names[]={"add2","sub2","mul2","fma2","mul2->add2","mul2->sub2","c-mul2","div3","(a+b)-a","fma2(4,b,10a)"};
  for(int i=0;i<10;i++) printf("%s mismatches: %u\n",names[i],hb[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
