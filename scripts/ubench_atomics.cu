// Microbenchmark: what bounds the scatter-add?  Random 256-byte rows in a 512 MB table (or an L2-resident 32 MB one):
// RED.32 / RED.v2 / RED.v4 (/ .v8 if supported) vs plain ST.v4 vs LD.v4, all with the 8-lanes-per-row mapping of libxdr.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)

template<int MODE> __global__ void k(float* tab, const int* idx, int n, float* sink){
  int g = (blockIdx.x*blockDim.x + threadIdx.x)/8, sub = threadIdx.x & 7;
  int ng = gridDim.x*blockDim.x/8;
  float acc=0.f;
  for(int i=g;i<n;i+=ng){
    float* row = tab + (size_t)idx[i]*64;
    #pragma unroll
    for(int v=0;v<2;++v){
      float4* p = reinterpret_cast<float4*>(row) + sub + v*8;
      if(MODE==0){ asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};"::"l"(p),"f"(1.f),"f"(2.f),"f"(3.f),"f"(4.f):"memory"); }
      else if(MODE==1){ float* q=(float*)p; for(int e=0;e<4;++e) asm volatile("red.global.add.f32 [%0], %1;"::"l"(q+e),"f"(1.f):"memory"); }
      else if(MODE==2){ float* q=(float*)p; for(int e=0;e<2;++e) asm volatile("red.global.add.v2.f32 [%0], {%1,%2};"::"l"(q+2*e),"f"(1.f),"f"(2.f):"memory"); }
      else if(MODE==3){ *p = make_float4(1.f,2.f,3.f,4.f); }
      else if(MODE==4){ float4 x = __ldcg(p); acc += x.x+x.y+x.z+x.w; }
      else if(MODE==5){ float4 x = __ldcg(p); x.x+=1.f; x.y+=2.f; x.z+=3.f; x.w+=4.f; *p = x; }
    }
  }
  if(MODE==4 && acc==123.456f) *sink=acc;
}
#ifdef TRY_V8
__global__ void k8(float* tab, const int* idx, int n){
  int g = (blockIdx.x*blockDim.x + threadIdx.x)/8, sub = threadIdx.x & 7;
  int ng = gridDim.x*blockDim.x/8;
  for(int i=g;i<n;i+=ng){
    float* p = tab + (size_t)idx[i]*64 + sub*8;
    asm volatile("red.global.add.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"::"l"(p),"f"(1.f),"f"(2.f),"f"(3.f),"f"(4.f),"f"(1.f),"f"(2.f),"f"(3.f),"f"(4.f):"memory");
  }
}
#endif
int main(){
  const int N = 3*8192*64;  // rows touched per launch (= 64 BPR steps' scatter)
  for(int big=1; big>=0; --big){
    size_t rows = big ? (2u<<20) : (128u<<10);  // 512 MB or 32 MB
    float* tab; CK(cudaMalloc(&tab, rows*256)); CK(cudaMemset(tab,0,rows*256));
    std::vector<int> h(N); std::mt19937 rng(1); for(auto& x:h) x = rng()%rows;
    int* idx; CK(cudaMalloc(&idx,N*4)); CK(cudaMemcpy(idx,h.data(),N*4,cudaMemcpyHostToDevice));
    float* sink; CK(cudaMalloc(&sink,4));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[]={"red.v4","red.f32 x4","red.v2 x2","st.v4","ld.v4","ld+st.v4 (non-atomic RMW)"};
    for(int mode=0; mode<6; ++mode){
      float best=1e9;
      for(int rep=0;rep<5;++rep){
        cudaEventRecord(e0);
        switch(mode){
          case 0: k<0><<<148*8,256>>>(tab,idx,N,sink); break; case 1: k<1><<<148*8,256>>>(tab,idx,N,sink); break;
          case 2: k<2><<<148*8,256>>>(tab,idx,N,sink); break; case 3: k<3><<<148*8,256>>>(tab,idx,N,sink); break;
          case 4: k<4><<<148*8,256>>>(tab,idx,N,sink); break; case 5: k<5><<<148*8,256>>>(tab,idx,N,sink); break; }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms;
      }
      printf("%s table  %-28s %8.1f us  -> %7.1f GB/s of row bytes, %6.2f us per 8192x3 rows\n", big?"512MB":" 32MB", names[mode], best*1e3, (double)N*256/best/1e6, best*1e3/64);
    }
#ifdef TRY_V8
    { float best=1e9; for(int rep=0;rep<5;++rep){ cudaEventRecord(e0); k8<<<148*8,256>>>(tab,idx,N); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
      printf("%s table  %-28s %8.1f us  -> %7.1f GB/s of row bytes, %6.2f us per 8192x3 rows\n", big?"512MB":" 32MB", "red.v8", best*1e3, (double)N*256/best/1e6, best*1e3/64); }
#endif
    cudaFree(tab); cudaFree(idx); cudaFree(sink);
  }
  return 0;
}
