// TMEM -> register load throughput (tcgen05.ld 32x32b.xN) as a function of the number of warps; sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include "../h-net-mamba-asr_b200/csrc/umma.cuh"
using namespace hnb;
template <int X>
__global__ void __launch_bounds__(512, 1) k(long long* out, int reps, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) umma::tmem_alloc(&slot, 512);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t t = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long c0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // 4 loads in flight, then one wait
      const uint32_t col = (uint32_t)(((warp >> 2) * 128 + i * 32) & 511);
      if (X == 32) umma::tmem_ld32(t + col, v);
      else if (X == 16) umma::tmem_ld16(t + col, v);
      else umma::tmem_ld8(t + col, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < X; ++j) acc += v[j];
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (threadIdx.x == 0) out[0] = c1 - c0;
  if (acc == 123.456f) *sink = acc;
  umma::tc_fence_before(); __syncthreads();
  if (warp == 0) umma::tmem_dealloc(slot, 512);
}
template <int X> void run(int warps, long long* d, float* sink) {
  const int reps = 200;
  k<X><<<1, warps * 32>>>(d, reps, sink);
  k<X><<<1, warps * 32>>>(d, reps, sink);
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * reps * 4 * 32 * X * 4;
  printf("x%-2d warps %2d: %8lld cycles, %.1f B/clk, %.0f cycles per ld+wait\n", X, warps, h, bytes / h, (double)h / (reps * 4));
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  for (int w : {1, 4, 8, 16}) { run<8>(w, d, s); run<16>(w, d, s); run<32>(w, d, s); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
