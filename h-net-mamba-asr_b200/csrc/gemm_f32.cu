// Exact fp32 GEMM on CUDA cores for the fp32 (decode / parity) path.  Tensor-core projections live in
// gemm_tcgen05.cu; this kernel exists because bf16/tf32 operands cannot meet the 1e-3 fp32 tolerance
// through 20 stacked layers, nor the 1e-4 boundary-probability margin of the router.
#include "common.cuh"

namespace hnb {

constexpr int GM = 64, GN = 64, GK = 16;

// C[m,n] = sum_k A(m,k) B(n,k);  A(m,k) = transA ? A[k*lda+m] : A[m*lda+k];  B(n,k) = transB ? B[k*ldb+n] : B[n*ldb+k]
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, long long lda, int transA, const float* __restrict__ Bm, long long ldb,
             int transB, int M, int N, int K, const float* __restrict__ bias, const float* __restrict__ R,
             long long ldr, float* __restrict__ Cm, long long ldc, int accumulate) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int i = tid; i < GM * GK; i += 256) {
      int m, k;
      if (transA) { m = i % GM; k = i / GM; } else { k = i % GK; m = i / GK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = transA ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
      As[k][m] = v;
    }
    for (int i = tid; i < GN * GK; i += 256) {
      int n, k;
      if (transB) { n = i % GN; k = i / GN; } else { k = i % GK; n = i / GK; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) v = transB ? Bm[(long long)gk * ldb + gn] : Bm[(long long)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (R) v += R[(long long)m * ldr + n];
      float* c = Cm + (long long)m * ldc + n;
      *c = accumulate ? (*c + v) : v;
    }
  }
}

}  // namespace hnb

using namespace hnb;

extern "C" int hnb_gemm_f32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                            int M, int N, int K, const float* bias, const float* R, long long ldr, float* C,
                            long long ldc, int accumulate, void* stream) {
  HNB_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_f32: bad arguments");
  dim3 grid(cdiv(N, GN), cdiv(M, GM));
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, transA, B, ldb, transB, M, N, K, bias, R, ldr, C, ldc,
                                                       accumulate);
  HNB_LAUNCH_CHECK("gemm_f32");
  return HNB_OK;
}
