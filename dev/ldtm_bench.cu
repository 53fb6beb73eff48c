// Microbenchmark: tcgen05.ld throughput per SM for several shapes / warp counts (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../multimodal-embedding_b200/csrc/sm100_ptx.cuh"
using namespace mmx;

template <int NREG> __device__ __forceinline__ void ld32x32(uint32_t taddr, uint32_t* v);
#define LD_ASM(N, ...) 
__device__ __forceinline__ void ld_x16(uint32_t t, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
   : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(t) : "memory");
}
__device__ __forceinline__ void ld_x64(uint32_t t, uint32_t (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
   : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]),"=r"(v[32]),"=r"(v[33]),"=r"(v[34]),"=r"(v[35]),"=r"(v[36]),"=r"(v[37]),"=r"(v[38]),"=r"(v[39]),"=r"(v[40]),"=r"(v[41]),"=r"(v[42]),"=r"(v[43]),"=r"(v[44]),"=r"(v[45]),"=r"(v[46]),"=r"(v[47]),"=r"(v[48]),"=r"(v[49]),"=r"(v[50]),"=r"(v[51]),"=r"(v[52]),"=r"(v[53]),"=r"(v[54]),"=r"(v[55]),"=r"(v[56]),"=r"(v[57]),"=r"(v[58]),"=r"(v[59]),"=r"(v[60]),"=r"(v[61]),"=r"(v[62]),"=r"(v[63]) : "r"(t) : "memory");
}
// 16x256b.x8: each warp reads 16 lanes x 8*256b... (x8 -> 32 regs/thread)
__device__ __forceinline__ void ld_16x256_x8(uint32_t t, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
   : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(t) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int nwarps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tbase;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      if (MODE == 0) { uint32_t v[32]; for (int c = 0; c < 8; ++c) { tmem_ld_32x32(t + c * 32 + (warp >> 2) * 256, v); tmem_ld_wait(); acc += v[0] ^ v[31]; } }
      if (MODE == 1) { uint32_t v[16]; for (int c = 0; c < 16; ++c) { ld_x16(t + c * 16 + (warp >> 2) * 256, v); tmem_ld_wait(); acc += v[0] ^ v[15]; } }
      if (MODE == 2) { uint32_t v[64]; for (int c = 0; c < 4; ++c) { ld_x64(t + c * 64 + (warp >> 2) * 256, v); tmem_ld_wait(); acc += v[0] ^ v[63]; } }
      if (MODE == 3) { uint32_t v[32]; uint32_t w[32]; for (int c = 0; c < 8; c += 2) { tmem_ld_32x32(t + c * 32 + (warp >> 2) * 256, v); tmem_ld_32x32(t + (c+1) * 32 + (warp >> 2) * 256, w); tmem_ld_wait(); acc += v[0] ^ w[31]; } }
      if (MODE == 4) { uint32_t v[32]; for (int c = 0; c < 8; ++c) { ld_16x256_x8(tbase + ((uint32_t)((warp & 3) * 32) << 16) + c * 32 + (warp >> 2) * 256, v); tmem_ld_wait(); acc += v[0] ^ v[31]; } }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  long long* out; uint32_t* sink; cudaMalloc(&out, 8 * 256); cudaMalloc(&sink, 4);
  const char* names[] = {"32x32b.x32 (wait each)", "32x32b.x16 (wait each)", "32x32b.x64 (wait each)", "32x32b.x32 x2 in flight", "16x256b.x8 (wait each)"};
  int iters = 2000;
  for (int mode = 0; mode < 5; ++mode) for (int nw : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) bench<0><<<148, 512>>>(nw, iters, out, sink);
      if (mode == 1) bench<1><<<148, 512>>>(nw, iters, out, sink);
      if (mode == 2) bench<2><<<148, 512>>>(nw, iters, out, sink);
      if (mode == 3) bench<3><<<148, 512>>>(nw, iters, out, sink);
      if (mode == 4) bench<4><<<148, 512>>>(nw, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)nw * iters * 32.0 * 256 * 4;   // each warp-iter reads 32 lanes x 256 cols x 4B
    printf("%-28s warps=%2d  cycles=%10lld  bytes/cycle/SM=%7.1f  cycles per 4KB warp-load=%6.1f\n", names[mode], nw, h, bytes / h, (double)h / (iters * 8.0));
  }
  return 0;
}
