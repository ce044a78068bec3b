// FixedPoolChunker (src/dcasr/models/fixed_pool.py:54-106), the reference's parameter-free control for the learned
// chunker: masked mean over fixed windows of `stride` frames, and the broadcast of each window vector back over its
// frames.  Window of frame t: w(t) = min(t / stride, M - 1) -- the reference clamps the tail of a padded batch into the
// last window (:80-83) -- so window j owns frames [j*stride, (j+1)*stride) and, for j = M - 1, everything up to L.
// Two bandwidth kernels serve forward and backward of both halves:
//   window_reduce    z[b,j]   = sum_{t in window j} m[b,t] x[b,t]  (/ max(cnt[b,j], 1))      chunk fwd | dechunk bwd
//   window_broadcast out[b,t] = m[b,t] / max(cnt[b,w],1) * z[b,w(t)]  (+ resid[b,t])         dechunk fwd | chunk bwd
// One warp per output row, lanes stride over D (rows are read as whole contiguous vectors); sums in fp32 (:84-86).
#include "common.cuh"

namespace hnb {
namespace {

constexpr int FP_WARPS = 8;

template <typename TX, typename TZ>
__global__ void __launch_bounds__(FP_WARPS * 32)
window_reduce_kernel(const TX* __restrict__ x, const uint8_t* __restrict__ mask, int B, int L, int D, int M, int stride,
                     int normalize, TZ* __restrict__ z, float* __restrict__ cnt_out) {
  const int lane = threadIdx.x & 31;
  const long long slot = (long long)blockIdx.x * FP_WARPS + (threadIdx.x >> 5);
  if (slot >= (long long)B * M) return;
  const int bi = (int)(slot / M), j = (int)(slot % M);
  const int t0 = min(j * stride, L), t1 = (j == M - 1) ? L : min((j + 1) * stride, L);
  const uint8_t* mrow = mask ? mask + (long long)bi * L : nullptr;
  float cnt = 0.f;
  for (int t = t0; t < t1; ++t) cnt += mrow ? (mrow[t] ? 1.f : 0.f) : 1.f;
  const TX* xr = x + (long long)bi * L * D;
  for (int c = lane; c < D; c += 32) {
    float acc = 0.f;
    for (int t = t0; t < t1; ++t)
      if (!mrow || mrow[t]) acc += to_f(xr[(long long)t * D + c]);
    z[slot * D + c] = from_f<TZ>(normalize ? acc / fmaxf(cnt, 1.f) : acc);
  }
  if (cnt_out && lane == 0) cnt_out[slot] = cnt;
}

template <typename TZ, typename TY>
__global__ void __launch_bounds__(FP_WARPS * 32)
window_broadcast_kernel(const TZ* __restrict__ z, const uint8_t* __restrict__ mask, const float* __restrict__ cnt,
                        const TY* __restrict__ resid, int B, int L, int D, int M, int stride, TY* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * FP_WARPS + (threadIdx.x >> 5);
  if (tok >= (long long)B * L) return;
  const int bi = (int)(tok / L), t = (int)(tok % L);
  const int w = min(t / stride, M - 1);
  const bool on = !mask || mask[tok];
  const float denom = cnt ? fmaxf(cnt[(long long)bi * M + w], 1.f) : 1.f;
  const TZ* src = z + ((long long)bi * M + w) * D;
  for (int c = lane; c < D; c += 32) {
    float v = on ? to_f(src[c]) / denom : 0.f;
    if (resid) v += to_f(resid[tok * D + c]);
    out[tok * D + c] = from_f<TY>(v);
  }
}

}  // namespace
}  // namespace hnb

using namespace hnb;

extern "C" int hnb_window_reduce(const void* x, int x_dtype, const uint8_t* mask, int B, int L, int D, int M, int stride,
                                 int normalize, void* z, int z_dtype, float* cnt, void* stream) {
  HNB_CHECK_ARG(x && z, "window_reduce: null pointer");
  HNB_CHECK_ARG(B > 0 && L > 0 && D > 0 && M >= 1 && stride >= 1, "window_reduce: bad sizes");
  const int grid = cdiv((long long)B * M, FP_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  HNB_DISPATCH_DTYPE(x_dtype, TX, HNB_DISPATCH_DTYPE(z_dtype, TZ, (window_reduce_kernel<TX, TZ><<<grid, FP_WARPS * 32, 0, st>>>(
      (const TX*)x, mask, B, L, D, M, stride, normalize, (TZ*)z, cnt))));
  HNB_LAUNCH_CHECK("window_reduce");
  return HNB_OK;
}

extern "C" int hnb_window_broadcast(const void* z, int z_dtype, const uint8_t* mask, const float* cnt, const void* resid,
                                    int B, int L, int D, int M, int stride, void* out, int out_dtype, void* stream) {
  HNB_CHECK_ARG(z && out, "window_broadcast: null pointer");
  HNB_CHECK_ARG(B > 0 && L > 0 && D > 0 && M >= 1 && stride >= 1, "window_broadcast: bad sizes");
  const int grid = cdiv((long long)B * L, FP_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  HNB_DISPATCH_DTYPE(z_dtype, TZ, HNB_DISPATCH_DTYPE(out_dtype, TY, (window_broadcast_kernel<TZ, TY><<<grid, FP_WARPS * 32, 0, st>>>(
      (const TZ*)z, mask, cnt, (const TY*)resid, B, L, D, M, stride, (TY*)out))));
  HNB_LAUNCH_CHECK("window_broadcast");
  return HNB_OK;
}
