// Causal depthwise conv1d (k=4) + SiLU over the xBC columns of zxbcdt, fused with dt = softplus(dt+bias)
// and with the length-aware sequence reversal of the backward-direction mixer: rows are read through
// scan_to_nat(), results are written in scan order, so no gather ever materialises a reversed copy.
//
// Bandwidth kernels: every thread owns 4 channels (one 8-byte bf16 / 16-byte fp32 vector) and a short run
// of scan positions; ALL of its row loads (run + 3 halo rows) are issued before any arithmetic so that each
// thread keeps 15-25 independent vector requests in flight (the loop-carried 4-tap window would otherwise
// serialise load -> use -> load).  Halo rows are re-read by the neighbouring tile from L2.
#include "common.cuh"

namespace hnb {

constexpr int CONV_TS = 16;     // scan positions per thread, forward
constexpr int CONV_TSB = 8;     // scan positions per thread and tile, backward
constexpr int CONV_GX = 8;      // backward: tile groups per (row, direction); each block strides over its tiles

// 4 channels per thread: 16-byte vectors for fp32, 8-byte vectors for bf16 (8 channels per thread would need
// ~250 registers for the taps, the sliding windows and the parameter-gradient accumulators: 1-2 CTAs per SM).
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  using raw_t = uint4;
  static __device__ __forceinline__ raw_t zero() { return make_uint4(0, 0, 0, 0); }
  static __device__ __forceinline__ raw_t ldg(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void st(float* p, const raw_t& r) { *reinterpret_cast<uint4*>(p) = r; }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  static __device__ __forceinline__ raw_t pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 4;
  using raw_t = uint2;
  static __device__ __forceinline__ raw_t zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ raw_t ldg(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const raw_t& r) { *reinterpret_cast<uint2*>(p) = r; }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ raw_t pack(const float* v) {
    raw_t r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __floats2bfloat162_rn(v[0], v[1]); h[1] = __floats2bfloat162_rn(v[2], v[3]);
    return r;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const T* __restrict__ zx, long long ldz, long long dstride, const int* __restrict__ lengths,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ dt_bias,
                int ndir, int B, int L, int di, int N, int H, T* __restrict__ xconv, float* __restrict__ dt_out) {
  constexpr int VN = V16<T>::N;
  const int dir = blockIdx.z, bi = blockIdx.y, s0 = blockIdx.x * CONV_TS;
  const int C = di + 2 * N;
  const long long T_ = (long long)B * L;
  const int len = lengths ? lengths[bi] : L;
  const long long xoff = (long long)dir * dstride + di;
  const long long doff = (long long)dir * dstride + di + C;
  const T* rowbase = zx + (long long)bi * L * ldz;
  const int s1 = min(s0 + CONV_TS, L);

  for (int c = threadIdx.x * VN; c < C; c += blockDim.x * VN) {
    typename V16<T>::raw_t raw[CONV_TS + 3];
#pragma unroll
    for (int k = 0; k < CONV_TS + 3; ++k) {
      const int s = s0 - 3 + k;
      raw[k] = (s >= 0 && s < L) ? V16<T>::ldg(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c)
                                 : V16<T>::zero();
    }
    float w[VN][4], bias[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4));
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = __ldg(conv_b + (long long)dir * C + c + i);
    }
    float win[3][VN], cur[VN], o[VN];
    V16<T>::unpack(raw[0], win[0]); V16<T>::unpack(raw[1], win[1]); V16<T>::unpack(raw[2], win[2]);
#pragma unroll
    for (int k = 0; k < CONV_TS; ++k) {
      const int s = s0 + k;
      V16<T>::unpack(raw[k + 3], cur);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
        o[i] = silu_f(pre);
        win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
      }
      if (s < s1) V16<T>::st(xconv + ((long long)dir * T_ + (long long)bi * L + s) * C + c, V16<T>::pack(o));
    }
  }
  for (int idx = threadIdx.x; idx < (s1 - s0) * H; idx += blockDim.x) {
    const int s = s0 + idx / H, h = idx % H;
    const float raw = to_f(rowbase[(long long)scan_to_nat(dir, s, len) * ldz + doff + h]);
    dt_out[((long long)dir * T_ + (long long)bi * L + s) * H + h] = softplus_f(raw + dt_bias[dir * H + h]);
  }
}

// backward.  dout[s] for channel c comes from dxc (c < di) or dBC (c >= di), both of the activation dtype; the
// pre-activation is recomputed from zxbcdt.  d input[s'] = sum_j w[j] dpre[s'+3-j];  dw[j] = sum_s dpre[s] in[s-3+j].
template <typename T>
__global__ void __launch_bounds__(256, 2)
conv_bwd_kernel(const T* __restrict__ zx, const T* __restrict__ dxc, long long ldz, long long dstride,
                const T* __restrict__ dBC, const float* __restrict__ ddt, const int* __restrict__ lengths,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ dt_bias,
                int ndir, int B, int L, int di, int N, int H, T* __restrict__ dzx, float* __restrict__ dconv_w,
                float* __restrict__ dconv_b, float* __restrict__ ddt_bias) {
  constexpr int VN = V16<T>::N;
  constexpr int TS = CONV_TSB;
  __shared__ float s_dtb[64];
  const int dir = blockIdx.z, bi = blockIdx.y;
  const int C = di + 2 * N;
  const long long T_ = (long long)B * L;
  const int len = lengths ? lengths[bi] : L;
  const long long xoff = (long long)dir * dstride + di;
  const long long doff = (long long)dir * dstride + di + C;
  const T* rowbase = zx + (long long)bi * L * ldz;
  T* drowbase = dzx + (long long)bi * L * ldz;
  const long long sbase = (long long)dir * T_ + (long long)bi * L;
  const int ntiles = (L + TS - 1) / TS;
  if (threadIdx.x < 64) s_dtb[threadIdx.x] = 0.f;
  __syncthreads();

  for (int c = threadIdx.x * VN; c < C; c += blockDim.x * VN) {
    float w[VN][4], bias[VN], gw[VN][4], gb[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4));
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = __ldg(conv_b + (long long)dir * C + c + i);
      gw[i][0] = gw[i][1] = gw[i][2] = gw[i][3] = 0.f; gb[i] = 0.f;
    }
    const bool is_x = c < di;
    const T* gsrc = is_x ? dxc + sbase * di + c : dBC + sbase * (2 * N) + (c - di);
    const long long gld = is_x ? di : 2 * N;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int s0 = tile * TS;
      const int s1 = min(s0 + TS, L);
      typename V16<T>::raw_t xr[TS + 6], gr[TS + 3];
#pragma unroll
      for (int k = 0; k < TS + 6; ++k) {                               // inputs at s0-3 .. s0+TS+2
        const int s = s0 - 3 + k;
        xr[k] = (s >= 0 && s < L) ? V16<T>::ldg(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c)
                                  : V16<T>::zero();
      }
#pragma unroll
      for (int k = 0; k < TS + 3; ++k) {                               // upstream grads at s0 .. s0+TS+2
        const int s = s0 + k;
        gr[k] = (s < L) ? V16<T>::ldg(gsrc + (long long)s * gld) : V16<T>::zero();
      }
      float win[3][VN], dp[3][VN];
      V16<T>::unpack(xr[0], win[0]); V16<T>::unpack(xr[1], win[1]); V16<T>::unpack(xr[2], win[2]);
#pragma unroll
      for (int i = 0; i < VN; ++i) { dp[0][i] = 0.f; dp[1][i] = 0.f; dp[2][i] = 0.f; }
#pragma unroll
      for (int k = 0; k < TS + 3; ++k) {
        const int s = s0 + k;
        float cur[VN], g[VN], dcur[VN];
        V16<T>::unpack(xr[k + 3], cur);
        V16<T>::unpack(gr[k], g);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
          const float sg = sigmoid_f(pre);
          dcur[i] = g[i] * sg * (1.f + pre * (1.f - sg));               // g is zero beyond L
          if (k < TS && s < s1) {                                       // parameter grads: the tile's own positions
            gw[i][0] += dcur[i] * win[0][i]; gw[i][1] += dcur[i] * win[1][i];
            gw[i][2] += dcur[i] * win[2][i]; gw[i][3] += dcur[i] * cur[i];
            gb[i] += dcur[i];
          }
        }
        if (k >= 3) {                                                   // d input at sp = s - 3 is complete
          const int sp = s - 3;
          if (sp < s1) {
            float o[VN];
#pragma unroll
            for (int i = 0; i < VN; ++i)
              o[i] = w[i][3] * dp[0][i] + w[i][2] * dp[1][i] + w[i][1] * dp[2][i] + w[i][0] * dcur[i];
            V16<T>::st(drowbase + (long long)scan_to_nat(dir, sp, len) * ldz + xoff + c, V16<T>::pack(o));
          }
        }
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
          dp[0][i] = dp[1][i]; dp[1][i] = dp[2][i]; dp[2][i] = dcur[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      float* gwp = dconv_w + ((long long)dir * C + c + i) * 4;
      atomicAdd(gwp + 0, gw[i][0]); atomicAdd(gwp + 1, gw[i][1]);
      atomicAdd(gwp + 2, gw[i][2]); atomicAdd(gwp + 3, gw[i][3]);
      atomicAdd(dconv_b + (long long)dir * C + c + i, gb[i]);
    }
  }
  // dt: d raw = ddt * sigmoid(raw + bias)
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int s0 = tile * TS;
    const int s1 = min(s0 + TS, L);
    for (int idx = threadIdx.x; idx < (s1 - s0) * H; idx += blockDim.x) {
      const int s = s0 + idx / H, h = idx % H;
      const long long nat = (long long)scan_to_nat(dir, s, len) * ldz + doff + h;
      const float raw = to_f(rowbase[nat]);
      const float g = ddt[(sbase + s) * H + h] * sigmoid_f(raw + dt_bias[dir * H + h]);
      drowbase[nat] = from_f<T>(g);
      atomicAdd(&s_dtb[h], g);
    }
  }
  __syncthreads();
  if (threadIdx.x < H) atomicAdd(ddt_bias + dir * H + threadIdx.x, s_dtb[threadIdx.x]);
}

}  // namespace hnb

using namespace hnb;

static int conv_check(const char* who, int dtype, long long ldz, long long dstride, int ndir, int B, int L, int di, int N,
                      int H) {
  const int vn = 4;
  (void)dtype;
  if (!(ndir >= 1 && ndir <= 2 && B > 0 && L > 0 && di > 0 && N > 0 && H > 0 && H <= 64)) {
    set_error("%s: bad sizes", who); return HNB_ERR_INVALID_ARG;
  }
  if (di % vn || (2 * N) % vn || ldz % vn || dstride % vn || dstride < 2LL * di + 2 * N + H || ldz < ndir * dstride) {
    set_error("%s: di, 2N, ldz, dstride must be multiples of %d, dstride >= 2di+2N+H, ldz >= ndir*dstride", who, vn);
    return HNB_ERR_INVALID_ARG;
  }
  return HNB_OK;
}

static int conv_threads(int C, int vn) {
  int t = (C / vn + 31) / 32 * 32;
  return t > 256 ? 256 : (t < 32 ? 32 : t);
}

extern "C" int hnb_conv_fwd(const void* zxbcdt, int dtype, long long ldz, long long dstride, const int32_t* lengths,
                            const float* conv_w, const float* conv_b, const float* dt_bias, int ndir, int B, int L,
                            int di, int N, int H, void* xconv, float* dt, void* stream) {
  HNB_CHECK_ARG(zxbcdt && conv_w && conv_b && dt_bias && xconv && dt, "conv_fwd: null pointer");
  int rc = conv_check("conv_fwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  dim3 grid(cdiv(L, CONV_TS), B, ndir);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = di + 2 * N;
  if (dtype == HNB_BF16)
    conv_fwd_kernel<__nv_bfloat16><<<grid, conv_threads(C, 4), 0, st>>>((const __nv_bfloat16*)zxbcdt, ldz, dstride,
        lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (__nv_bfloat16*)xconv, dt);
  else if (dtype == HNB_F32)
    conv_fwd_kernel<float><<<grid, conv_threads(C, 4), 0, st>>>((const float*)zxbcdt, ldz, dstride, lengths, conv_w,
        conv_b, dt_bias, ndir, B, L, di, N, H, (float*)xconv, dt);
  else { set_error("conv_fwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_fwd");
  return HNB_OK;
}

extern "C" int hnb_conv_bwd(const void* zxbcdt, const void* dxc, int dtype, long long ldz, long long dstride,
                            const void* dBC, const float* ddt, const int32_t* lengths, const float* conv_w,
                            const float* conv_b, const float* dt_bias, int ndir, int B, int L, int di, int N, int H,
                            void* dzxbcdt, float* dconv_w, float* dconv_b, float* ddt_bias, void* stream) {
  HNB_CHECK_ARG(zxbcdt && dxc && dBC && ddt && conv_w && conv_b && dt_bias && dzxbcdt && dconv_w && dconv_b && ddt_bias,
                "conv_bwd: null pointer");
  int rc = conv_check("conv_bwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  const int ntiles = cdiv(L, CONV_TSB);
  dim3 grid(ntiles < CONV_GX ? ntiles : CONV_GX, B, ndir);
  cudaStream_t st = (cudaStream_t)stream;
  const int C = di + 2 * N;
  if (dtype == HNB_BF16)
    conv_bwd_kernel<__nv_bfloat16><<<grid, conv_threads(C, 4), 0, st>>>((const __nv_bfloat16*)zxbcdt,
        (const __nv_bfloat16*)dxc, ldz, dstride, (const __nv_bfloat16*)dBC, ddt, lengths, conv_w, conv_b, dt_bias, ndir,
        B, L, di, N, H, (__nv_bfloat16*)dzxbcdt, dconv_w, dconv_b, ddt_bias);
  else if (dtype == HNB_F32)
    conv_bwd_kernel<float><<<grid, conv_threads(C, 4), 0, st>>>((const float*)zxbcdt, (const float*)dxc, ldz, dstride,
        (const float*)dBC, ddt, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (float*)dzxbcdt, dconv_w,
        dconv_b, ddt_bias);
  else { set_error("conv_bwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_bwd");
  return HNB_OK;
}
