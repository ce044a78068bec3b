// CTC head after the encoder: log-softmax + CTC loss over the logits of `Linear d_outer -> V+1`,
// forward and backward, without ever materialising the fp32 logits or the [B, L, V+1] fp32 log-probabilities.
//
// Reference (/root/reference/src/dcasr/decoders/ctc.py:100-115):
//     lp   = F.log_softmax(proj(features).float(), -1)                               # fp32 [B, L, V+1], two round trips
//     loss = F.ctc_loss(lp.transpose(0, 1), targets, feat_lengths, target_lengths, blank, reduction, zero_infinity=True)
// Here, on the logits as the projection GEMM wrote them (bf16 under autocast, fp32 otherwise):
//     hnb_ctc_lse        lse[b, t] = logsumexp_c logits[b, t, c]  (+ the per-frame argmax for greedy decoding)    one read
//     hnb_ctc_alpha_beta alpha / beta recursions in log space, one CTA per (utterance, direction); the only log-
//                        probabilities they touch are the 2U+1 gathered per frame: lp = logit - lse
//     hnb_ctc_grad       dlogits[b, t, c] = g_b (softmax - occupancy) in the logits' dtype: one read, one write
// The recursions follow the published algorithm (Graves et al. 2006) in PyTorch's convention: alpha and beta both include
// the emission at t, so the occupancy of class c at t is exp(logsum_{s: l'_s = c}(alpha + beta) + nll - lp).
#include <cfloat>

#include "common.cuh"

namespace hnb {
namespace {

constexpr float NEG_INF = -INFINITY;

__device__ __forceinline__ float log_add(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == NEG_INF) return NEG_INF;
  return m + log1pf(__expf(-fabsf(a - b)));
}
__device__ __forceinline__ float log_add3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == NEG_INF) return NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// ---- lse + argmax: one warp per frame -------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
ctc_lse_kernel(const T* __restrict__ logits, long long rows, int V1, long long ldl, float* __restrict__ lse,
               int* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* x = logits + row * ldl;
  float m = NEG_INF;
  int am = 0;
  for (int c = lane; c < V1; c += 32) {
    const float v = to_f(x[c]);
    if (v > m) { m = v; am = c; }                       // first maximum wins inside a lane (ascending c)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }       // ties -> lowest class id, as torch.argmax
  }
  float s = 0.f;
  for (int c = lane; c < V1; c += 32) s += __expf(to_f(x[c]) - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    lse[row] = m + __logf(s);
    if (argmax) argmax[row] = am;
  }
}

// ---- alpha / beta: CTA (b, dir), thread s = extended-label position ---------------------------------
// alpha[b, t, s], beta[b, t, s] fp32, s < S = 2 U + 1 (U = padded target width).  nll[b] = -log p(l | x).
template <typename T>
__global__ void ctc_alpha_beta_kernel(const T* __restrict__ logits, const float* __restrict__ lse,
                                      const long long* __restrict__ targets, const long long* __restrict__ feat_lens,
                                      const long long* __restrict__ tgt_lens, int Tm, int V1, long long ldl, int U, int blank,
                                      float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ nll) {
  extern __shared__ float sh[];                          // two buffers of S + 2 (guards at both ends)
  const int b = blockIdx.x >> 1, dir = blockIdx.x & 1;
  const int S = 2 * U + 1, s = threadIdx.x;
  const int Tb = (int)min((long long)Tm, max(0LL, feat_lens[b]));
  const int Ub = (int)min((long long)U, max(0LL, tgt_lens[b]));
  const int Sb = 2 * Ub + 1;
  const bool live = s < Sb;
  const long long* tg = targets + (long long)b * U;
  int lab = blank, lab2 = -1;                            // l'_s and the label two positions towards the recursion's source
  if (live && (s & 1)) {
    lab = (int)tg[s >> 1];
    if (dir == 0) lab2 = s >= 3 ? (int)tg[(s >> 1) - 1] : -1;
    else lab2 = (s >> 1) + 1 < Ub ? (int)tg[(s >> 1) + 1] : -1;
  }
  const bool skip_ok = (s & 1) && lab2 >= 0 && lab2 != lab;           // the s -/+ 2 transition
  const bool lab_ok = lab >= 0 && lab < V1;
  float* buf0 = sh + 2;
  float* buf1 = sh + (S + 4) + 2;
  for (int i = threadIdx.x; i < 2 * (S + 4); i += blockDim.x) sh[i] = NEG_INF;
  __syncthreads();
  float* out = (dir == 0 ? alpha : beta) + (long long)b * Tm * S;
  const T* lg = logits + (long long)b * Tm * ldl;
  const float* ls = lse + (long long)b * Tm;
  if (Tb == 0) {
    if (dir == 0 && s == 0) nll[b] = Ub == 0 ? 0.f : INFINITY;
    return;
  }
  const int t0 = dir == 0 ? 0 : Tb - 1, step = dir == 0 ? 1 : -1;
  float lp_next = (live && lab_ok) ? to_f(lg[(long long)t0 * ldl + lab]) - ls[t0] : NEG_INF;
  float* cur = buf0;
  float* prv = buf1;
  for (int i = 0, t = t0; i < Tb; ++i, t += step) {
    const float lp = lp_next;
    if (i + 1 < Tb && live && lab_ok) {                  // the next frame's emission, fetched under this frame's arithmetic
      const int tn = t + step;
      lp_next = to_f(lg[(long long)tn * ldl + lab]) - ls[tn];
    }
    float v = NEG_INF;
    if (live) {
      if (i == 0) {
        if (dir == 0) v = s <= 1 ? lp : NEG_INF;
        else v = s >= Sb - 2 ? lp : NEG_INF;
      } else {
        const float a0 = prv[s];
        const float a1 = dir == 0 ? prv[s - 1] : (s + 1 < Sb ? prv[s + 1] : NEG_INF);
        const float a2 = skip_ok ? (dir == 0 ? prv[s - 2] : prv[s + 2]) : NEG_INF;
        v = log_add3(a0, a1, a2) + lp;
        if (!(v > NEG_INF)) v = NEG_INF;                  // -inf + x stays -inf; NaN never enters the lattice
      }
    }
    if (s < S) {
      cur[s] = v;
      out[(long long)t * S + s] = v;
    }
    __syncthreads();
    float* tmp = cur; cur = prv; prv = tmp;
  }
  if (dir == 0 && s == 0) {
    const float l = log_add(prv[Sb - 1], Sb >= 2 ? prv[Sb - 2] : NEG_INF);
    nll[b] = -l;
  }
}

// ---- gradient: one CTA per frame ----------------------------------------------------------------------
// dlogits[b, t, c] = g[b] * (softmax[c] - exp(logsum_{s: l'_s = c}(alpha + beta)[t, s] + nll[b] - lp[c])),  zero for t >= T_b,
// for an infeasible utterance (nll = inf, zero_infinity) and where g[b] = 0.
template <typename T>
__global__ void __launch_bounds__(128)
ctc_grad_kernel(const T* __restrict__ logits, const float* __restrict__ lse, const float* __restrict__ alpha,
                const float* __restrict__ beta, const long long* __restrict__ targets, const long long* __restrict__ feat_lens,
                const long long* __restrict__ tgt_lens, const float* __restrict__ nll, const float* __restrict__ gscale, int Tm,
                int V1, long long ldl, int U, int blank, T* __restrict__ dlogits, long long ldd) {
  extern __shared__ float occ[];                         // [V1] linear-domain occupancy sums relative to the frame's maximum
  __shared__ float red[4];
  const int b = blockIdx.x / Tm, t = blockIdx.x - b * Tm;
  const int Tb = (int)min((long long)Tm, max(0LL, feat_lens[b]));
  const int Ub = (int)min((long long)U, max(0LL, tgt_lens[b]));
  const int S = 2 * U + 1, Sb = 2 * Ub + 1;
  T* dl = dlogits + ((long long)b * Tm + t) * ldd;
  const float g = gscale[b], nl = nll[b];
  if (t >= Tb || g == 0.f || !(nl < INFINITY)) {
    for (int c = threadIdx.x; c < V1; c += blockDim.x) dl[c] = from_f<T>(0.f);
    return;
  }
  const float* a = alpha + ((long long)b * Tm + t) * S;
  const float* be = beta + ((long long)b * Tm + t) * S;
  const long long* tg = targets + (long long)b * U;
  float m = NEG_INF;
  for (int s = threadIdx.x; s < Sb; s += blockDim.x) m = fmaxf(m, a[s] + be[s]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  for (int c = threadIdx.x; c < V1; c += blockDim.x) occ[c] = 0.f;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  if (m > NEG_INF) {
    for (int s = threadIdx.x; s < Sb; s += blockDim.x) {
      const int lab = (s & 1) ? (int)tg[s >> 1] : blank;
      if (lab >= 0 && lab < V1) atomicAdd(&occ[lab], __expf(a[s] + be[s] - m));
    }
  }
  __syncthreads();
  const T* x = logits + ((long long)b * Tm + t) * ldl;
  const float l = lse[(long long)b * Tm + t];
  for (int c = threadIdx.x; c < V1; c += blockDim.x) {
    const float lp = to_f(x[c]) - l;
    const float o = occ[c];
    const float w = o > 0.f ? __expf(__logf(o) + m + nl - lp) : 0.f;
    dl[c] = from_f<T>(g * (__expf(lp) - w));
  }
}

// ---- column sums (bias gradients): out[c] += sum_r x[r, c] --------------------------------------------
// A block owns ROWS_PER_BLOCK rows and every column; thread c-strided columns are read coalesced row by row, the block's
// partial sums leave through one atomic per column.
constexpr int CS_ROWS = 128;
template <typename T>
__global__ void __launch_bounds__(256)
col_sum_kernel(const T* __restrict__ x, long long rows, int cols, long long ldx, float* __restrict__ out) {
  const long long r0 = (long long)blockIdx.x * CS_ROWS;
  const long long r1 = min(rows, r0 + CS_ROWS);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += to_f(x[r * ldx + c]);
    atomicAdd(out + c, acc);
  }
}

}  // namespace
}  // namespace hnb

using namespace hnb;

extern "C" int hnb_col_sum(const void* x, int dtype, long long rows, int cols, long long ldx, float* out, void* stream) {
  HNB_CHECK_ARG(x && out && rows >= 0 && cols > 0 && ldx >= cols, "col_sum: bad arguments");
  if (rows == 0) return HNB_OK;
  HNB_DISPATCH_DTYPE(dtype, T, (col_sum_kernel<T><<<cdiv(rows, CS_ROWS), 256, 0, (cudaStream_t)stream>>>((const T*)x, rows, cols,
                                                                                                       ldx, out)));
  HNB_LAUNCH_CHECK("col_sum");
  return HNB_OK;
}

extern "C" int hnb_ctc_lse(const void* logits, int dtype, long long rows, int V1, long long ldl, float* lse, int* argmax,
                           void* stream) {
  HNB_CHECK_ARG(logits && lse && rows >= 0 && V1 > 0 && ldl >= V1, "ctc_lse: bad arguments");
  if (rows == 0) return HNB_OK;
  const int wpb = 8;
  HNB_DISPATCH_DTYPE(dtype, T, (ctc_lse_kernel<T><<<cdiv(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
                                   (const T*)logits, rows, V1, ldl, lse, argmax)));
  HNB_LAUNCH_CHECK("ctc_lse");
  return HNB_OK;
}

extern "C" int hnb_ctc_alpha_beta(const void* logits, int dtype, const float* lse, const long long* targets,
                                  const long long* feat_lens, const long long* tgt_lens, int B, int T, int V1, long long ldl,
                                  int U, int blank, float* alpha, float* beta, float* nll, void* stream) {
  HNB_CHECK_ARG(logits && lse && feat_lens && tgt_lens && alpha && beta && nll && (targets || U == 0), "ctc_alpha_beta: null pointer");
  HNB_CHECK_ARG(B > 0 && T > 0 && V1 > 0 && U >= 0 && blank >= 0 && blank < V1, "ctc_alpha_beta: bad sizes");
  const int S = 2 * U + 1;
  HNB_CHECK_ARG(S <= 1024, "ctc_alpha_beta: at most 511 target tokens per utterance (got %d)", U);
  const int threads = (S + 31) / 32 * 32;
  const size_t smem = 2 * (size_t)(S + 4) * sizeof(float);
  HNB_DISPATCH_DTYPE(dtype, Ty, (ctc_alpha_beta_kernel<Ty><<<2 * B, threads, smem, (cudaStream_t)stream>>>(
                                    (const Ty*)logits, lse, targets, feat_lens, tgt_lens, T, V1, ldl, U, blank, alpha, beta, nll)));
  HNB_LAUNCH_CHECK("ctc_alpha_beta");
  return HNB_OK;
}

extern "C" int hnb_ctc_grad(const void* logits, int dtype, const float* lse, const float* alpha, const float* beta,
                            const long long* targets, const long long* feat_lens, const long long* tgt_lens, const float* nll,
                            const float* gscale, int B, int T, int V1, long long ldl, int U, int blank, void* dlogits,
                            long long ldd, void* stream) {
  HNB_CHECK_ARG(logits && lse && alpha && beta && feat_lens && tgt_lens && nll && gscale && dlogits && (targets || U == 0),
                "ctc_grad: null pointer");
  HNB_CHECK_ARG(B > 0 && T > 0 && V1 > 0 && U >= 0 && blank >= 0 && blank < V1 && ldd >= V1, "ctc_grad: bad sizes");
  HNB_CHECK_ARG((size_t)V1 * sizeof(float) <= 48 * 1024, "ctc_grad: at most 12288 classes");
  HNB_DISPATCH_DTYPE(dtype, Ty, (ctc_grad_kernel<Ty><<<B * T, 128, (size_t)V1 * sizeof(float), (cudaStream_t)stream>>>(
                                    (const Ty*)logits, lse, alpha, beta, targets, feat_lens, tgt_lens, nll, gscale, T, V1, ldl, U,
                                    blank, (Ty*)dlogits, ldd)));
  HNB_LAUNCH_CHECK("ctc_grad");
  return HNB_OK;
}
