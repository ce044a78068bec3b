// fp32 GEMMs of the fp32 (decode / parity) path.  A single bf16 or tf32 product cannot meet the 1e-3 fp32 tolerance through
// 20 stacked layers, nor the 1e-4 boundary-probability margin of the router, so there are two kernels:
//   * hnb_gemm_f32     exact fp32 on CUDA cores (small problems, and the yard-stick of the other one);
//   * hnb_gemm_f32_tc  fp32 operands split into bf16 PIECES  a = a0 + a1 + a2  (8 + 8 + 8 significant bits), the product
//                      rebuilt from the piece products that matter -- a0 b0 + a0 b1 + a1 b0 + a1 b1 + a0 b2 + a2 b0 (dropped:
//                      terms below 2^-24 |a||b|) -- on the tcgen05 GEMM with fp32 accumulation in TMEM, over the staged
//                      operands A' = [a0 | a1 | a1 | a0 | a2 | a0],  B' = [b1 | b0 | b1 | b2 | b0 | b0].   fp32-class accuracy (measured
//                      against the exact kernel in tests/test_gpu_mamba.py) at ~6 bf16 GEMMs' cost, which is still ~10 x
//                      the CUDA-core kernel's speed: the production decode path (reference tasks/decode_task.py:123-151
//                      runs the encoder in fp32) spent 63 % of its time in hnb_gemm_f32.
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

// ---- piece splitting ------------------------------------------------------------------------------------
// x [rows, cols] fp32 (row stride ldx) -> out: NT segments along K.  k_rows = 0: K is the column axis, segment s occupies
// columns [s*seg, s*seg + cols) of out (row stride ldo; the pad up to seg is zero-filled); k_rows = 1: K is the row axis,
// segment s occupies rows [s*seg, s*seg + rows).  piece[s] selects which bf16 piece (0, 1, 2) goes into segment s.
struct PieceMap { int piece[6]; };
__global__ void __launch_bounds__(256)
split_pieces_kernel(const float* __restrict__ x, long long ldx, int rows, int cols, int k_rows, int seg, int nt, PieceMap pm,
                    __nv_bfloat16* __restrict__ out, long long ldo) {
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int r = blockIdx.y;
  const int cols_pad = k_rows ? cols : seg;                   // columns this launch covers (incl. the zero pad of a K segment)
  if (c4 >= cols_pad) return;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (r < rows && c4 + j < cols) ? x[(long long)r * ldx + c4 + j] : 0.f;
  __nv_bfloat16 pc[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float rem = v[j];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      pc[q][j] = __float2bfloat16_rn(rem);
      rem -= __bfloat162float(pc[q][j]);
    }
  }
  for (int s = 0; s < nt; ++s) {
    const int q = pm.piece[s];
    __nv_bfloat16* o = k_rows ? out + ((long long)s * seg + r) * ldo + c4 : out + (long long)r * ldo + (long long)s * seg + c4;
    if (c4 + 3 < (k_rows ? (int)ldo : seg)) {
      *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(&pc[q][0]);
    } else {
      for (int j = 0; j < 4 && c4 + j < (k_rows ? (int)ldo : seg); ++j) o[j] = pc[q][j];
    }
  }
}

inline int up8(int v) { return (v + 7) / 8 * 8; }

}  // namespace hnb

using namespace hnb;

extern "C" long long hnb_gemm_f32_tc_ws_bytes(int M, int N, int K) {
  // A': non-transposed [M, 6 K8] or transposed [6 K8, M8]; both sizes are covered by the larger expression
  const long long k8 = up8(K), a = (long long)up8(M) * 6 * k8, b = (long long)up8(N) * 6 * k8;
  return (a + b) * 2 + 512;
}

// C = op(A) op(B) (+ bias) (+ R), operand conventions of hnb_gemm_f32; ws: hnb_gemm_f32_tc_ws_bytes(M, N, K) bytes, 256-byte aligned.
extern "C" int hnb_gemm_f32_tc(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, int M, int N,
                               int K, const float* bias, const float* R, long long ldr, float* C, long long ldc, void* ws,
                               void* stream) {
  HNB_CHECK_ARG(A && B && C && ws && M > 0 && N > 0 && K > 0, "gemm_f32_tc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int k8 = up8(K), nt = 6, kp = nt * k8;
  __nv_bfloat16* Ap = reinterpret_cast<__nv_bfloat16*>(ws);
  const long long a_elems = (long long)up8(M) * kp;
  __nv_bfloat16* Bp = Ap + (a_elems + 127) / 128 * 128;
  // segment order: the five small products first, a0 b0 last (see below)
  const PieceMap pa = {{0, 1, 1, 0, 2, 0}}, pb = {{1, 0, 1, 2, 0, 0}};
  // A(m, k): transA ? A[k * lda + m] : A[m * lda + k]
  long long lda_p, ldb_p;
  if (!transA) {
    lda_p = kp;
    split_pieces_kernel<<<dim3(cdiv(k8, 1024), M), 256, 0, st>>>(A, lda, M, K, 0, k8, nt, pa, Ap, lda_p);
  } else {
    lda_p = up8(M);
    split_pieces_kernel<<<dim3(cdiv(M, 1024), k8), 256, 0, st>>>(A, lda, K, M, 1, k8, nt, pa, Ap, lda_p);
  }
  HNB_LAUNCH_CHECK("gemm_f32_tc split A");
  if (!transB) {
    ldb_p = kp;
    split_pieces_kernel<<<dim3(cdiv(k8, 1024), N), 256, 0, st>>>(B, ldb, N, K, 0, k8, nt, pb, Bp, ldb_p);
  } else {
    ldb_p = up8(N);
    split_pieces_kernel<<<dim3(cdiv(N, 1024), k8), 256, 0, st>>>(B, ldb, K, N, 1, k8, nt, pb, Bp, ldb_p);
  }
  HNB_LAUNCH_CHECK("gemm_f32_tc split B");
  // Two passes over the same staged operands: the five small piece products (each <= 2^-8 of a0 b0) are summed first, then the
  // a0 b0 pass adds its K terms on top.  The tensor core's fp32 accumulation loses ~2^-24 of the running sum per step; in
  // one 6 K pass every one of those steps carries the full-size sum (measured 4e-6 relative error at K = 384), this way only
  // K of them do, which is the error of an fp32 accumulation of length K -- what the exact kernel has.
  const long long small = 5LL * k8;
  int rc = hnb_gemm_bf16(Ap, lda_p, transA, Bp, ldb_p, transB, M, N, (int)small, bias, R, ldr, C, ldc, HNB_F32, 1, stream);
  if (rc != HNB_OK) return rc;
  const __nv_bfloat16* A0 = transA ? Ap + small * lda_p : Ap + small;
  const __nv_bfloat16* B0 = transB ? Bp + small * ldb_p : Bp + small;
  // long K: the a0 b0 pass in slices of <= 512, each added onto C through the epilogue's residual input (C = slice + C, round to
  // nearest), so that the per-step loss above is bounded by the slice length (K = 5003 in one slice: 5.5e-6; the exact kernel:
  // 1.3e-6; sliced: 4.8e-7) and the result stays bit-reproducible (no atomics).
  for (int k0 = 0; k0 < k8; k0 += 512) {
    const int kk = k8 - k0 < 512 ? k8 - k0 : 512;
    const __nv_bfloat16* As = transA ? A0 + (long long)k0 * lda_p : A0 + k0;
    const __nv_bfloat16* Bs = transB ? B0 + (long long)k0 * ldb_p : B0 + k0;
    rc = hnb_gemm_bf16(As, lda_p, transA, Bs, ldb_p, transB, M, N, kk, nullptr, C, ldc, C, ldc, HNB_F32, 1, stream);
    if (rc != HNB_OK) return rc;
  }
  return HNB_OK;
}

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
