// Causal depthwise conv1d (k=4) + SiLU over the xBC columns of zxbcdt, fused with dt = softplus(dt+bias)
// and with the length-aware sequence reversal of the backward-direction mixer: rows are read through
// scan_to_nat(), results are written in scan order, so no gather ever materialises a reversed copy.
// Each thread owns VN channels and slides a 4-row register window along the tile (smem-free halo reuse;
// neighbouring tiles' 3 halo rows come from L2).
#include "common.cuh"

namespace hnb {

constexpr int CONV_TS = 16;     // scan positions per forward tile
constexpr int CONV_TSB = 32;    // scan positions per backward tile
constexpr int CONV_THREADS = 128;

template <typename T, int VN>
__global__ void __launch_bounds__(CONV_THREADS)
conv_fwd_kernel(const T* __restrict__ zx, long long ldz, long long dstride, const int* __restrict__ lengths,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ dt_bias,
                int ndir, int B, int L, int di, int N, int H, T* __restrict__ xconv, float* __restrict__ dt_out) {
  const int dir = blockIdx.z, bi = blockIdx.y, s0 = blockIdx.x * CONV_TS;
  const int C = di + 2 * N;
  const long long T_ = (long long)B * L;
  const int len = lengths ? lengths[bi] : L;
  const long long xoff = (long long)dir * dstride + di;
  const long long doff = (long long)dir * dstride + di + C;
  const T* rowbase = zx + (long long)bi * L * ldz;
  const int s1 = min(s0 + CONV_TS, L);

  for (int c = threadIdx.x * VN; c < C; c += CONV_THREADS * VN) {
    float w[VN][4], bias[VN], win[3][VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4);
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = conv_b[(long long)dir * C + c + i];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int s = s0 - 3 + k;
#pragma unroll
      for (int i = 0; i < VN; ++i) win[k][i] = 0.f;
      if (s >= 0) ldv<T, VN>(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c, win[k]);
    }
    for (int s = s0; s < s1; ++s) {
      float cur[VN], o[VN];
      ldv<T, VN>(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c, cur);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
        o[i] = silu_f(pre);
        win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
      }
      stv<T, VN>(xconv + ((long long)dir * T_ + (long long)bi * L + s) * C + c, o);
    }
  }
  for (int idx = threadIdx.x; idx < (s1 - s0) * H; idx += CONV_THREADS) {
    const int s = s0 + idx / H, h = idx % H;
    const float raw = to_f(rowbase[(long long)scan_to_nat(dir, s, len) * ldz + doff + h]);
    dt_out[((long long)dir * T_ + (long long)bi * L + s) * H + h] = softplus_f(raw + dt_bias[dir * H + h]);
  }
}

// backward.  dout[s] for channel c comes from dxc (c < di) or dBC (c >= di); the pre-activation is
// recomputed from zxbcdt.  d input[s'] = sum_j w[j] dpre[s'+3-j];  dw[j] = sum_s dpre[s] in[s-3+j].
template <typename T, int VN>
__global__ void __launch_bounds__(CONV_THREADS)
conv_bwd_kernel(const T* __restrict__ zx, const T* __restrict__ dxc, long long ldz, long long dstride, const float* __restrict__ dBC,
                const float* __restrict__ ddt, const int* __restrict__ lengths, const float* __restrict__ conv_w,
                const float* __restrict__ conv_b, const float* __restrict__ dt_bias, int ndir, int B, int L, int di,
                int N, int H, T* __restrict__ dzx, float* __restrict__ dconv_w, float* __restrict__ dconv_b,
                float* __restrict__ ddt_bias) {
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
  const int ntiles = (L + CONV_TSB - 1) / CONV_TSB;
  if (threadIdx.x < 64) s_dtb[threadIdx.x] = 0.f;
  __syncthreads();

  for (int c = threadIdx.x * VN; c < C; c += CONV_THREADS * VN) {
    float w[VN][4], bias[VN], gw[VN][4], gb[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4);
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = conv_b[(long long)dir * C + c + i];
      gw[i][0] = gw[i][1] = gw[i][2] = gw[i][3] = 0.f; gb[i] = 0.f;
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int s0 = tile * CONV_TSB;
      const int s1 = min(s0 + CONV_TSB, L);
      float win[3][VN], dp[3][VN];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int s = s0 - 3 + k;
#pragma unroll
        for (int i = 0; i < VN; ++i) { win[k][i] = 0.f; dp[k][i] = 0.f; }
        if (s >= 0) ldv<T, VN>(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c, win[k]);
      }
      // s runs 3 positions past the tile so that d input of the last tile rows sees its future dpre
      for (int s = s0; s < s1 + 3; ++s) {
        float cur[VN], dcur[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) { cur[i] = 0.f; dcur[i] = 0.f; }
        if (s < L) {
          ldv<T, VN>(rowbase + (long long)scan_to_nat(dir, s, len) * ldz + xoff + c, cur);
          float g[VN];
          if (c < di) ldv<T, VN>(dxc + (sbase + s) * di + c, g);
          else ldv<float, VN>(dBC + (sbase + s) * (2 * N) + (c - di), g);
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
            const float sg = sigmoid_f(pre);
            dcur[i] = g[i] * sg * (1.f + pre * (1.f - sg));
            if (s < s1) {                                           // parameter grads: own positions only
              gw[i][0] += dcur[i] * win[0][i]; gw[i][1] += dcur[i] * win[1][i];
              gw[i][2] += dcur[i] * win[2][i]; gw[i][3] += dcur[i] * cur[i];
              gb[i] += dcur[i];
            }
          }
        }
        const int sp = s - 3;                                       // input position whose gradient is complete
        if (sp >= s0 && sp < s1) {
          float o[VN];
#pragma unroll
          for (int i = 0; i < VN; ++i)
            o[i] = w[i][3] * dp[0][i] + w[i][2] * dp[1][i] + w[i][1] * dp[2][i] + w[i][0] * dcur[i];
          stv<T, VN>(drowbase + (long long)scan_to_nat(dir, sp, len) * ldz + xoff + c, o);
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
    const int s0 = tile * CONV_TSB;
    const int s1 = min(s0 + CONV_TSB, L);
    for (int idx = threadIdx.x; idx < (s1 - s0) * H; idx += CONV_THREADS) {
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

static int conv_check(const char* who, int dtype, long long ldz, long long dstride, int ndir, int B, int L, int di, int N, int H) {
  const int vn = dtype == HNB_BF16 ? 8 : 4;
  if (!(ndir >= 1 && ndir <= 2 && B > 0 && L > 0 && di > 0 && N > 0 && H > 0 && H <= 64)) {
    set_error("%s: bad sizes", who); return HNB_ERR_INVALID_ARG;
  }
  if (di % vn || (2 * N) % vn || ldz % vn || dstride % vn || dstride < 2LL * di + 2 * N + H || ldz < ndir * dstride) {
    set_error("%s: di, 2N, ldz, dstride must be multiples of %d, dstride >= 2di+2N+H, ldz >= ndir*dstride", who, vn);
    return HNB_ERR_INVALID_ARG;
  }
  return HNB_OK;
}

extern "C" int hnb_conv_fwd(const void* zxbcdt, int dtype, long long ldz, long long dstride, const int32_t* lengths, const float* conv_w,
                            const float* conv_b, const float* dt_bias, int ndir, int B, int L, int di, int N, int H,
                            void* xconv, float* dt, void* stream) {
  HNB_CHECK_ARG(zxbcdt && conv_w && conv_b && dt_bias && xconv && dt, "conv_fwd: null pointer");
  int rc = conv_check("conv_fwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  dim3 grid(cdiv(L, CONV_TS), B, ndir);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HNB_BF16)
    conv_fwd_kernel<__nv_bfloat16, 8><<<grid, CONV_THREADS, 0, st>>>((const __nv_bfloat16*)zxbcdt, ldz, dstride, lengths, conv_w,
        conv_b, dt_bias, ndir, B, L, di, N, H, (__nv_bfloat16*)xconv, dt);
  else if (dtype == HNB_F32)
    conv_fwd_kernel<float, 4><<<grid, CONV_THREADS, 0, st>>>((const float*)zxbcdt, ldz, dstride, lengths, conv_w, conv_b,
        dt_bias, ndir, B, L, di, N, H, (float*)xconv, dt);
  else { set_error("conv_fwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_fwd");
  return HNB_OK;
}

extern "C" int hnb_conv_bwd(const void* zxbcdt, const void* dxc, int dtype, long long ldz, long long dstride, const float* dBC,
                            const float* ddt, const int32_t* lengths, const float* conv_w, const float* conv_b,
                            const float* dt_bias, int ndir, int B, int L, int di, int N, int H, void* dzxbcdt,
                            float* dconv_w, float* dconv_b, float* ddt_bias, void* stream) {
  HNB_CHECK_ARG(zxbcdt && dxc && dBC && ddt && conv_w && conv_b && dt_bias && dzxbcdt && dconv_w && dconv_b && ddt_bias,
                "conv_bwd: null pointer");
  int rc = conv_check("conv_bwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  const int ntiles = cdiv(L, CONV_TSB);
  dim3 grid(ntiles < 4 ? ntiles : 4, B, ndir);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HNB_BF16)
    conv_bwd_kernel<__nv_bfloat16, 8><<<grid, CONV_THREADS, 0, st>>>((const __nv_bfloat16*)zxbcdt,
        (const __nv_bfloat16*)dxc, ldz, dstride, dBC, ddt, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H,
        (__nv_bfloat16*)dzxbcdt, dconv_w, dconv_b, ddt_bias);
  else if (dtype == HNB_F32)
    conv_bwd_kernel<float, 4><<<grid, CONV_THREADS, 0, st>>>((const float*)zxbcdt, (const float*)dxc, ldz, dstride, dBC, ddt,
        lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (float*)dzxbcdt, dconv_w, dconv_b, ddt_bias);
  else { set_error("conv_bwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_bwd");
  return HNB_OK;
}
