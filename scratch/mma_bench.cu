// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) by operand major-ness and N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I h-net-mamba-asr_b200/csrc -I include scratch/mma_bench.cu -o scratch/mma_bench -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
#include "umma.cuh"
using namespace hnb;
constexpr int HALF = 128 * 128;
template <int MODE, int NMMA>
__global__ void __launch_bounds__(128, 1) k(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem; uint8_t* sB = smem + 2 * HALF;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * HALF);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 4 * HALF / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
  if (threadIdx.x < 32) umma::tmem_alloc(slot, 128);
  umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        const int kb = i & 7;
        const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
        uint64_t da, db; uint32_t id;
        if (MODE == 0) { da = umma::make_smem_desc(umma::smem_u32(sA) + o, 16, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024); id = umma::make_idesc_bf16(128, 128, 0, 0); }
        if (MODE == 1) { da = umma::make_smem_desc(umma::smem_u32(sA) + o, 16, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, 1024, 1024); id = umma::make_idesc_bf16(128, 64, 0, 1); }
        if (MODE == 2) { da = umma::make_smem_desc(umma::smem_u32(sA) + kb * 2048, HALF, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, 1024, 1024); id = umma::make_idesc_bf16(128, 64, 1, 1); }
        if (MODE == 3) { da = umma::make_smem_desc(umma::smem_u32(sA) + o, 16, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024); id = umma::make_idesc_bf16(128, 128, 0, 1); }
        if (MODE == 4) { da = umma::make_smem_desc(umma::smem_u32(sA) + kb * 2048, HALF, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024); id = umma::make_idesc_bf16(128, 128, 1, 1); }
        if (MODE == 5) { da = umma::make_smem_desc(umma::smem_u32(sA) + o, 16, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + (kb & 3) * 32, 16, 1024); id = umma::make_idesc_bf16(128, 64, 0, 0); }
        if (MODE == 6) { da = umma::make_smem_desc(umma::smem_u32(sA) + kb * 2048, HALF, 1024); db = umma::make_smem_desc(umma::smem_u32(sB) + (kb & 3) * 32, 16, 1024); id = umma::make_idesc_bf16(128, 64, 1, 0); }
        umma::mma_bf16_ss(tmem, da, db, id, i > 0);
      }
      const long long t1 = clock64();
      umma::mma_commit(bar);
      umma::mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      if (blockIdx.x == 0 && rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  umma::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 128);
}
template <int MODE> void run(const char* name, long long* d) {
  const int smem = 4 * HALF + 1024;
  cudaFuncSetAttribute(k<MODE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<MODE, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h8[2], h40[2];
  k<MODE, 8><<<148, 128, smem>>>(d); cudaMemcpy(h8, d, 16, cudaMemcpyDeviceToHost);
  k<MODE, 40><<<148, 128, smem>>>(d); cudaMemcpy(h40, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-28s 8 MMAs: issue %5lld total %5lld | 40 MMAs: issue %5lld total %5lld | per MMA %.1f cyc  %s\n", name, h8[0], h8[1], h40[0],
         h40[1], (h40[1] - h8[1]) / 32.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<0>("K/K   N=128 (G, R)", d);
  run<5>("K/K   N=64", d);
  run<1>("K/MN  N=64  (Yo, du2)", d);
  run<6>("MN/K  N=64", d);
  run<2>("MN/MN N=64  (dS, du1)", d);
  run<3>("K/MN  N=128 (dC += W B)", d);
  run<4>("MN/MN N=128 (dB += W^T C)", d);
  return 0;
}
