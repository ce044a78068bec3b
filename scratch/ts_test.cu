// Unit check of tcgen05.mma with the A operand in TMEM (bf16 pairs packed per 32-bit column, lane = row).
//   D[128 x 64] = A[128 x 128] * B[64 x 128]^T     A: TMEM (written by tcgen05.st), B: smem K-major SWIZZLE_128B
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "umma.cuh"
using namespace hnb;
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                                     // 2 blocks [64 rows x 64 k]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 128; i += 128) {
    const int n = i / 128, kk = i % 128, blk = kk / 64, k6 = kk % 64;
    const uint32_t off = blk * 8192 + n * 128 + (((k6 >> 3) ^ (n & 7)) << 4) + (k6 & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = B[i];
  }
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
  if (warp == 0) umma::tmem_alloc(slot, 128);
  umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = *slot, t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c8 = 0; c8 < 8; ++c8) {                        // row tid: 128 bf16 -> 64 packed columns at TMEM 64..127
    uint32_t pk[8];
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 v = __halves2bfloat162(A[tid * 128 + 16 * c8 + 2 * j], A[tid * 128 + 16 * c8 + 2 * j + 1]);
      pk[j] = *reinterpret_cast<const uint32_t*>(&v);
    }
    tmem_st8(t_lane + 64 + 8 * c8, pk);
  }
  tmem_st_wait();
  umma::tc_fence_before(); __syncthreads();
  if (tid == 0) {
    umma::tc_fence_after();
    constexpr uint32_t id = umma::make_idesc_bf16(128, 64, 0, 0);
    for (int kb = 0; kb < 8; ++kb)
      mma_bf16_ts(tmem + 0, tmem + 64 + 8 * kb, umma::make_smem_desc(umma::smem_u32(sB) + (kb >> 2) * 8192 + (kb & 3) * 32, 16, 1024), id, kb > 0);
    umma::mma_commit(bar);
  }
  umma::mbar_wait(bar, 0);
  umma::tc_fence_after();
  float v[32];
  for (int j = 0; j < 2; ++j) {
    umma::tmem_ld32(t_lane + 32 * j, v);
    umma::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[tid * 64 + 32 * j + i] = v[i];
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}
int main() {
  std::vector<__nv_bfloat16> A(128 * 128), B(64 * 128);
  std::vector<float> Af(128 * 128), Bf(64 * 128), D(128 * 64);
  srand(1);
  for (int i = 0; i < 128 * 128; ++i) { A[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Af[i] = __bfloat162float(A[i]); }
  for (int i = 0; i < 64 * 128; ++i) { B[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Bf[i] = __bfloat162float(B[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1088);
  k<<<1, 128, 16384 + 1088>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
    double r = 0; for (int kk = 0; kk < 128; ++kk) r += (double)Af[m * 128 + kk] * Bf[n * 128 + kk];
    maxerr = fmax(maxerr, fabs(r - D[m * 64 + n]));
  }
  printf("ts-mma: %s, max abs err %.3e (expect < 1e-3)\n", cudaGetErrorString(e), maxerr);
  return 0;
}
