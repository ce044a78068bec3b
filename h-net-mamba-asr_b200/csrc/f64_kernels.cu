// float64 instantiations of the two chunk-stage ops whose reference tests run torch.autograd.gradcheck in double precision
// (reference tests/test_hnet_chunk.py:242-263 on DynamicChunker._ema, tests/test_fixed_pool.py:197-207 on the fixed-stride
// pool / broadcast).  The production dtypes are fp32 and bf16; these kernels exist so that the drop-in passes the reference's
// own test files unmodified, on tensors of a few hundred elements -- one thread per (row, channel), plain loops, no tuning.
//   hnb_ema_fwd_f64 / hnb_ema_bwd_f64          out_0 = x_0; out_t = pc_t x_t + (1 - pc_t) out_{t-1}, pc = clamp(P, c, 1 - c),
//                                              zero gradient to P outside [c, 1 - c]   (src/dcasr/models/hnet_chunk.py:226-248)
//   hnb_window_reduce_f64 / hnb_window_broadcast_f64   the fp64 twins of csrc/fixed_pool_kernels.cu
#include "common.cuh"

namespace hnb {
namespace {

__global__ void ema_fwd_f64_kernel(const double* __restrict__ x, const double* __restrict__ P, int B, int M, int D, double pcl,
                                   double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int b = (int)(i / D), c = (int)(i % D);
  const double* xr = x + (long long)b * M * D + c;
  double* orow = out + (long long)b * M * D + c;
  double prev = xr[0];
  orow[0] = prev;
  for (int t = 1; t < M; ++t) {
    const double pc = fmin(fmax(P[(long long)b * M + t], pcl), 1.0 - pcl);
    prev = pc * xr[(long long)t * D] + (1.0 - pc) * prev;
    orow[(long long)t * D] = prev;
  }
}

__global__ void ema_bwd_f64_kernel(const double* __restrict__ dout, const double* __restrict__ x, const double* __restrict__ out,
                                   const double* __restrict__ P, int B, int M, int D, double pcl, double* __restrict__ dx,
                                   double* __restrict__ dP) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int b = (int)(i / D), c = (int)(i % D);
  const long long base = (long long)b * M * D + c;
  double g = 0.0;                                           // d loss / d out_t including what flows back from t + 1
  for (int t = M - 1; t >= 1; --t) {
    g += dout[base + (long long)t * D];
    const double Pt = P[(long long)b * M + t];
    const double pc = fmin(fmax(Pt, pcl), 1.0 - pcl);
    dx[base + (long long)t * D] = pc * g;
    if (Pt >= pcl && Pt <= 1.0 - pcl) atomicAdd(dP + (long long)b * M + t, g * (x[base + (long long)t * D] - out[base + (long long)(t - 1) * D]));
    g *= (1.0 - pc);
  }
  dx[base] = g + dout[base];
}

__global__ void window_reduce_f64_kernel(const double* __restrict__ x, const uint8_t* __restrict__ mask, int B, int L, int D, int M,
                                         int stride, int normalize, double* __restrict__ z, float* __restrict__ cnt_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * M * D) return;
  const int c = (int)(i % D);
  const long long slot = i / D;
  const int bi = (int)(slot / M), j = (int)(slot % M);
  const int t0 = min(j * stride, L), t1 = (j == M - 1) ? L : min((j + 1) * stride, L);
  const uint8_t* mrow = mask ? mask + (long long)bi * L : nullptr;
  double acc = 0.0, cnt = 0.0;
  for (int t = t0; t < t1; ++t)
    if (!mrow || mrow[t]) { acc += x[((long long)bi * L + t) * D + c]; cnt += 1.0; }
  z[i] = normalize ? acc / fmax(cnt, 1.0) : acc;
  if (cnt_out && c == 0) cnt_out[slot] = (float)cnt;
}

__global__ void window_broadcast_f64_kernel(const double* __restrict__ z, const uint8_t* __restrict__ mask,
                                            const float* __restrict__ cnt, const double* __restrict__ resid, int B, int L, int D,
                                            int M, int stride, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * L * D) return;
  const int c = (int)(i % D);
  const long long tok = i / D;
  const int bi = (int)(tok / L), t = (int)(tok % L);
  const int w = min(t / stride, M - 1);
  const bool on = !mask || mask[tok];
  const double denom = cnt ? fmax((double)cnt[(long long)bi * M + w], 1.0) : 1.0;
  double v = on ? z[((long long)bi * M + w) * D + c] / denom : 0.0;
  if (resid) v += resid[i];
  out[i] = v;
}

}  // namespace
}  // namespace hnb

using namespace hnb;

extern "C" int hnb_ema_fwd_f64(const double* x, const double* P, int B, int M, int D, double p_clamp, double* out, void* stream) {
  HNB_CHECK_ARG(x && P && out && B > 0 && M > 0 && D > 0, "ema_fwd_f64: bad arguments");
  ema_fwd_f64_kernel<<<cdiv((long long)B * D, 128), 128, 0, (cudaStream_t)stream>>>(x, P, B, M, D, p_clamp, out);
  HNB_LAUNCH_CHECK("ema_fwd_f64");
  return HNB_OK;
}

extern "C" int hnb_ema_bwd_f64(const double* dout, const double* x, const double* out, const double* P, int B, int M, int D,
                               double p_clamp, double* dx, double* dP, void* stream) {
  HNB_CHECK_ARG(dout && x && out && P && dx && dP && B > 0 && M > 0 && D > 0, "ema_bwd_f64: bad arguments");
  ema_bwd_f64_kernel<<<cdiv((long long)B * D, 128), 128, 0, (cudaStream_t)stream>>>(dout, x, out, P, B, M, D, p_clamp, dx, dP);
  HNB_LAUNCH_CHECK("ema_bwd_f64");
  return HNB_OK;
}

extern "C" int hnb_window_reduce_f64(const double* x, const uint8_t* mask, int B, int L, int D, int M, int stride, int normalize,
                                     double* z, float* cnt, void* stream) {
  HNB_CHECK_ARG(x && z && B > 0 && L > 0 && D > 0 && M >= 1 && stride >= 1, "window_reduce_f64: bad arguments");
  window_reduce_f64_kernel<<<cdiv((long long)B * M * D, 128), 128, 0, (cudaStream_t)stream>>>(x, mask, B, L, D, M, stride, normalize,
                                                                                             z, cnt);
  HNB_LAUNCH_CHECK("window_reduce_f64");
  return HNB_OK;
}

extern "C" int hnb_window_broadcast_f64(const double* z, const uint8_t* mask, const float* cnt, const double* resid, int B, int L,
                                        int D, int M, int stride, double* out, void* stream) {
  HNB_CHECK_ARG(z && out && B > 0 && L > 0 && D > 0 && M >= 1 && stride >= 1, "window_broadcast_f64: bad arguments");
  window_broadcast_f64_kernel<<<cdiv((long long)B * L * D, 128), 128, 0, (cudaStream_t)stream>>>(z, mask, cnt, resid, B, L, D, M,
                                                                                                stride, out);
  HNB_LAUNCH_CHECK("window_broadcast_f64");
  return HNB_OK;
}
