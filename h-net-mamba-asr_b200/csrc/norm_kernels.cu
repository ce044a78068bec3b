// LayerNorm (pre-norm / final norm of MambaBlock / MambaStack) and the mixer's gated RMSNorm.
// One warp per row; rows are short (d_model 384..768, d_inner 768..1536) so a row lives in L1 and is
// re-read for the second moment instead of being held in a variable-size register array.
#include <algorithm>

#include "common.cuh"

namespace hnb {

constexpr int NORM_WARPS = 8;

// sigmoid of the gate: fp32 activations keep ex2 + rcp; bf16 activations take one MUFU (tanh.approx, abs. error
// ~5e-4, below the bf16 rounding of the values it multiplies)
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) { return sigmoid_f(x); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward
// ---------------------------------------------------------------------------------------------
template <typename TX, typename TY, int VN>
__global__ void __launch_bounds__(NORM_WARPS * 32)
layernorm_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     long long rows, int d, float eps, TY* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TX* xr = x + row * d;
  float s = 0.f;
  for (int c = lane * VN; c < d; c += 32 * VN) {
    float v[VN]; ldv<TX, VN>(xr + c, v);
#pragma unroll
    for (int i = 0; i < VN; ++i) s += v[i];
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
  for (int c = lane * VN; c < d; c += 32 * VN) {
    float v[VN]; ldv<TX, VN>(xr + c, v);
#pragma unroll
    for (int i = 0; i < VN; ++i) { const float t = v[i] - mean; q += t * t; }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  TY* yr = y + row * d;
  for (int c = lane * VN; c < d; c += 32 * VN) {
    float v[VN], g[VN], b[VN];
    ldv<TX, VN>(xr + c, v); ldv<float, VN>(gamma + c, g); ldv<float, VN>(beta + c, b);
#pragma unroll
    for (int i = 0; i < VN; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
    stv<TY, VN>(yr + c, v);
  }
}

// Register-resident variant for the widths of this model (d <= 768, multiple of 4): each row is read from global
// memory ONCE, two rows per warp with all their loads issued before the first reduction.  The generic kernel above
// walks a row three times and keeps a single 1.5 KB row in flight per warp -- too little to cover a DRAM round trip.
template <typename TX, typename TY, int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32)
layernorm_fwd_reg_kernel(const TX* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         long long rows, int d, float eps, TY* __restrict__ y, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out) {
  pdl_enter();
  constexpr int VN = 4, R = 2;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  float xv[R][NV][VN];
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
#pragma unroll
      for (int i = 0; i < VN; ++i) xv[u][k][i] = 0.f;
      if (c < d && row0 + u < rows) ldv<TX, VN>(x + (row0 + u) * d + c, xv[u][k]);
    }
  const float inv_d = 1.f / d;
#pragma unroll
  for (int u = 0; u < R; ++u) {
    if (row0 + u >= rows) break;                                     // warp-uniform
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int i = 0; i < VN; ++i) s += xv[u][k][i];
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < d) {
#pragma unroll
        for (int i = 0; i < VN; ++i) { const float t = xv[u][k][i] - mean; xv[u][k][i] = t; q = fmaf(t, t, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) { mean_out[row0 + u] = mean; rstd_out[row0 + u] = rstd; }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < d) {
        float g[VN], b[VN], o[VN];
        ldv<float, VN>(gamma + c, g); ldv<float, VN>(beta + c, b);
#pragma unroll
        for (int i = 0; i < VN; ++i) o[i] = fmaf(xv[u][k][i] * rstd, g[i], b[i]);
        stv<TY, VN>(y + (row0 + u) * d + c, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (+ residual gradient).  Persistent-style grid: each warp strides over rows and
// keeps its dgamma/dbeta columns in registers; one shared-memory reduction + atomics per block.
// ---------------------------------------------------------------------------------------------
template <typename TDY, typename TX, typename TDX, int VN, int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32, (NV * VN <= 16 ? 2 : 1))
layernorm_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const TDX* __restrict__ dres,
                     long long rows, int d, TDX* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  pdl_enter();
  extern __shared__ float sm[];                                     // [NORM_WARPS][2][d]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float ag[NV][VN], ab[NV][VN], gm[NV][VN];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * VN;
#pragma unroll
    for (int i = 0; i < VN; ++i) { ag[k][i] = 0.f; ab[k][i] = 0.f; gm[k][i] = 0.f; }
    if (c < d) ldv<float, VN>(gamma + c, gm[k]);
  }
  // two rows per iteration (d <= 384): both rows' loads are issued before either row's reduction, so a warp has twice the bytes
  // in flight (these 384..512-element rows are too short to hide a DRAM round trip otherwise)
  constexpr int R = NV * VN <= 12 ? 2 : 1;                          // wider rows: one at a time (register budget)
  const long long stride = (long long)gridDim.x * NORM_WARPS;
  for (long long row0 = (long long)blockIdx.x * NORM_WARPS + w; row0 < rows; row0 += R * stride) {
    float xv[R][NV][VN], g[R][NV][VN], rr[R][NV][VN];
    float mu[R], rs[R];
    bool have[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const long long row = row0 + u * stride;
      have[u] = row < rows;
      if (have[u]) {
        mu[u] = mean[row]; rs[u] = rstd[row];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c = (k * 32 + lane) * VN;
          if (c < d) {
            ldv<TX, VN>(x + row * d + c, xv[u][k]);
            ldv<TDY, VN>(dy + row * d + c, g[u][k]);
            if (dres) ldv<TDX, VN>(dres + row * d + c, rr[u][k]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      if (!have[u]) continue;                                        // warp-uniform
      const long long row = row0 + u * stride;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * VN;
        if (c < d) {
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const float xh = (xv[u][k][i] - mu[u]) * rs[u];
            xv[u][k][i] = xh;
            ag[k][i] += g[u][k][i] * xh;
            ab[k][i] += g[u][k][i];
            const float gg = g[u][k][i] * gm[k][i];
            g[u][k][i] = gg;
            s1 += gg; s2 += gg * xh;
          }
        }
      }
      s1 = warp_sum(s1) / d; s2 = warp_sum(s2) / d;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * VN;
        if (c < d) {
          float o[VN];
#pragma unroll
          for (int i = 0; i < VN; ++i)
            o[i] = rs[u] * (g[u][k][i] - s1 - xv[u][k][i] * s2) + (dres ? rr[u][k][i] : 0.f);
          stv<TDX, VN>(dx + row * d + c, o);
        }
      }
    }
  }
  // block reduction of the parameter gradients: every warp leaves its partial sums in its own [2][d] slice of shared memory, a
  // column is summed over the warps by one thread and leaves as a vector reduction (the first version: 6 k shared-memory float
  // atomics -- CAS loops -- and 2 d scalar global atomics per block, ~5 us of a 26 us kernel)
  {
    float* part = sm + (size_t)w * 2 * d;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < d) {
#pragma unroll
        for (int i = 0; i < VN; ++i) { part[c + i] = ag[k][i]; part[d + c + i] = ab[k][i]; }
      }
    }
  }
  __syncthreads();
  const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0;
  if (vec) {
    for (int c = threadIdx.x * 4; c < 2 * d; c += blockDim.x * 4) {
      float4 a = *reinterpret_cast<const float4*>(sm + c);
#pragma unroll
      for (int q = 1; q < NORM_WARPS; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(sm + (size_t)q * 2 * d + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      float* dst = c < d ? dgamma + c : dbeta + (c - d);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    }
  } else {
    for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < NORM_WARPS; ++q) a += sm[(size_t)q * 2 * d + c];
      atomicAdd(c < d ? dgamma + c : dbeta + (c - d), a);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// gated RMSNorm: warp per (direction, natural token).  y is read at the token's scan position, z from the
// natural row of zxbcdt; results land in natural order.  Both kernels are INSTRUCTION-bound (ncu: issue slots
// 60-70 % busy at 45-55 % of HBM peak), so the row is read from global memory once, kept packed in registers
// (NV vectors of VN elements per lane, fully unrolled) and every sigmoid is evaluated once.
// ---------------------------------------------------------------------------------------------
template <typename T, int VN> struct RawVec;
template <> struct RawVec<float, 1> { using t = float; };
template <> struct RawVec<float, 4> { using t = uint4; };
template <> struct RawVec<__nv_bfloat16, 1> { using t = __nv_bfloat16; };
template <> struct RawVec<__nv_bfloat16, 4> { using t = uint2; };
template <> struct RawVec<__nv_bfloat16, 8> { using t = uint4; };
template <typename T, int VN> __device__ __forceinline__ typename RawVec<T, VN>::t ld_raw(const T* p) {
  return *reinterpret_cast<const typename RawVec<T, VN>::t*>(p);
}
template <typename T, int VN> __device__ __forceinline__ void unpack_raw(const typename RawVec<T, VN>::t& r, float* v) {
  ldv<T, VN>(reinterpret_cast<const T*>(&r), v);          // register-to-register: the "load" folds into moves / shifts
}

template <typename T, int VN, int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32)
gated_norm_fwd_kernel(const T* __restrict__ y, const T* __restrict__ zx, long long ldz, long long dstride, const int* __restrict__ lengths,
                      const float* __restrict__ w, int ndir, int B, int L, int di, float eps, T* __restrict__ out,
                      float* __restrict__ rstd_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long T_ = (long long)B * L;
  const long long idx = (long long)blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  if (idx >= T_ * ndir) return;
  const int dir = (int)(idx / T_);
  const long long tok = idx % T_;
  const int bi = (int)(tok / L), t = (int)(tok % L);
  const int len = lengths ? min(max(lengths[bi], 0), L) : L;   // clamped like reverse_sequences (mamba_block.py:26)
  const int s = scan_to_nat(dir, t, len);                           // the map is an involution
  const T* yr = y + ((long long)dir * T_ + (long long)bi * L + s) * di;
  const T* zr = zx + tok * ldz + (long long)dir * dstride;
  typename RawVec<T, VN>::t ry[NV], rz[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * VN;
    if (c < di) { ry[k] = ld_raw<T, VN>(yr + c); rz[k] = ld_raw<T, VN>(zr + c); }
  }
  float g[NV][VN];
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * VN;
    if (c < di) {
      float a[VN], z[VN];
      unpack_raw<T, VN>(ry[k], a); unpack_raw<T, VN>(rz[k], z);
#pragma unroll
      for (int i = 0; i < VN; ++i) { g[k][i] = a[i] * z[i] * sigmoid_t<T>(z[i]); q = fmaf(g[k][i], g[k][i], q); }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / di + eps);
  if (lane == 0) rstd_out[(long long)dir * T_ + tok] = rs;
  T* o = out + tok * ((long long)ndir * di) + (long long)dir * di;
  const float* wr = w + (long long)dir * di;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = (k * 32 + lane) * VN;
    if (c < di) {
      float ww[VN];
      ldv<float, VN>(wr + c, ww);
#pragma unroll
      for (int i = 0; i < VN; ++i) g[k][i] *= rs * ww[i];
      stv<T, VN>(o + c, g[k]);
    }
  }
}

// backward: dout (natural) -> dy (scan order), dz (natural, into dzxbcdt), dw (accumulated).
template <typename T, int VN, int NV>
__global__ void __launch_bounds__(NORM_WARPS * 32, (NV * VN <= 32 ? 2 : 1))
gated_norm_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y, const T* __restrict__ zx, long long ldz,
                      long long dstride, const int* __restrict__ lengths, const float* __restrict__ w,
                      const float* __restrict__ rstd, int ndir, int B, int L, int di, T* __restrict__ dy,
                      T* __restrict__ dzx, float* __restrict__ dw) {
  pdl_enter();
  extern __shared__ float sm[];                                     // [NORM_WARPS][di]
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const long long T_ = (long long)B * L;
  const int dir = blockIdx.y;
  const float* wr = w + (long long)dir * di;
  float aw[NV][VN];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < VN; ++i) aw[k][i] = 0.f;
  for (long long tok = (long long)blockIdx.x * NORM_WARPS + wi; tok < T_; tok += (long long)gridDim.x * NORM_WARPS) {
    const int bi = (int)(tok / L), t = (int)(tok % L);
    const int len = lengths ? min(max(lengths[bi], 0), L) : L;   // clamped like reverse_sequences (mamba_block.py:26)
    const int s = scan_to_nat(dir, t, len);
    const long long yoff = ((long long)dir * T_ + (long long)bi * L + s) * di;
    const T* zr = zx + tok * ldz + (long long)dir * dstride;
    const T* gr = dout + tok * ((long long)ndir * di) + (long long)dir * di;
    const float rs = rstd[(long long)dir * T_ + tok];
    typename RawVec<T, VN>::t ry[NV], rz[NV], rg[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < di) { ry[k] = ld_raw<T, VN>(y + yoff + c); rz[k] = ld_raw<T, VN>(zr + c); rg[k] = ld_raw<T, VN>(gr + c); }
    }
    constexpr bool KEEP_SG = NV * VN <= 24;                         // wider rows: re-evaluate rather than spill
    float sg[KEEP_SG ? NV : 1][VN];                                 // sigmoid(z): one evaluation per element
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < di) {
        float yv[VN], zv[VN], go[VN], ww[VN];
        unpack_raw<T, VN>(ry[k], yv); unpack_raw<T, VN>(rz[k], zv); unpack_raw<T, VN>(rg[k], go);
        ldv<float, VN>(wr + c, ww);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float sgi = sigmoid_t<T>(zv[i]);
          if (KEEP_SG) sg[k][i] = sgi;
          s2 = fmaf(go[i] * ww[i], yv[i] * zv[i] * sgi, s2);
        }
      }
    }
    s2 = warp_sum(s2) * rs / di;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < di) {
        float yv[VN], zv[VN], go[VN], ww[VN], o1[VN], o2[VN];
        unpack_raw<T, VN>(ry[k], yv); unpack_raw<T, VN>(rz[k], zv); unpack_raw<T, VN>(rg[k], go);
        ldv<float, VN>(wr + c, ww);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float sgi = KEEP_SG ? sg[KEEP_SG ? k : 0][i] : sigmoid_t<T>(zv[i]);
          const float gh = yv[i] * zv[i] * sgi * rs;                 // normalised gated value
          aw[k][i] = fmaf(go[i], gh, aw[k][i]);
          const float dg = rs * (go[i] * ww[i] - gh * s2) * sgi;
          o1[i] = dg * zv[i];                                        // d y
          o2[i] = dg * yv[i] * fmaf(zv[i], 1.f - sgi, 1.f);          // d z
        }
        stv<T, VN>(dy + yoff + c, o1);
        stv<T, VN>(dzx + tok * ldz + (long long)dir * dstride + c, o2);
      }
    }
  }
  // dw: per-warp partial sums in shared memory, one thread per column quad sums them and issues one vector reduction
  {
    float* part = sm + (size_t)wi * di;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = (k * 32 + lane) * VN;
      if (c < di) {
#pragma unroll
        for (int i = 0; i < VN; ++i) part[c + i] = aw[k][i];
      }
    }
  }
  __syncthreads();
  float* dwr = dw + (long long)dir * di;
  if (di % 4 == 0 && (reinterpret_cast<uintptr_t>(dwr) & 15) == 0) {
    for (int c = threadIdx.x * 4; c < di; c += blockDim.x * 4) {
      float4 a = *reinterpret_cast<const float4*>(sm + c);
#pragma unroll
      for (int q = 1; q < NORM_WARPS; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(sm + (size_t)q * di + c);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dwr + c), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
    }
  } else {
    for (int c = threadIdx.x; c < di; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < NORM_WARPS; ++q) a += sm[(size_t)q * di + c];
      atomicAdd(dwr + c, a);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-step parameter packing of one mixer direction (replaces ~10 torch cat/stack/cast launches per block):
// in_proj.weight -> rows [dir*dstride, +dip) of Win (pad rows zeroed), out_proj.weight -> columns [dir*di, +di) of
// Wout, both cast to the activation dtype; the small fp32 vectors are copied into their [ndir, ...] stacks.
// ---------------------------------------------------------------------------------------------
constexpr int PACK_MAX_LAYERS = 24;
// per layer and direction: in_w, out_w, conv_w, conv_b, dt_bias, A_log, D, norm_w  (3 KB: passed by value as a kernel argument)
struct PackSrc { const float* p[PACK_MAX_LAYERS][2][8]; };

template <typename TW>
__global__ void __launch_bounds__(256)
pack_mixer_kernel(const PackSrc src, int dir0, int ndir, int d, int di, const FastDiv dd4, const FastDiv ddi4,
                  int N, int H, int dstride, TW* __restrict__ Win, TW* __restrict__ Wout, float* __restrict__ conv_w_o,
                  float* __restrict__ conv_b_o, float* __restrict__ dt_bias_o, float* __restrict__ A_log_o,
                  float* __restrict__ D_o, float* __restrict__ norm_w_o, long long layer_stride_bytes) {
  pdl_enter();
  const int dir = dir0 + blockIdx.y;                                // one launch packs gridDim.y directions of gridDim.z layers
  const int ly = blockIdx.z;
  {
    const long long off = (long long)ly * layer_stride_bytes;       // every output of layer ly sits `off` bytes after layer 0's
    Win = reinterpret_cast<TW*>(reinterpret_cast<uint8_t*>(Win) + off);
    Wout = reinterpret_cast<TW*>(reinterpret_cast<uint8_t*>(Wout) + off);
    conv_w_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(conv_w_o) + off);
    conv_b_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(conv_b_o) + off);
    dt_bias_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dt_bias_o) + off);
    A_log_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(A_log_o) + off);
    D_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(D_o) + off);
    norm_w_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(norm_w_o) + off);
  }
  const float* __restrict__ in_w = src.p[ly][blockIdx.y][0]; const float* __restrict__ out_w = src.p[ly][blockIdx.y][1];
  const float* __restrict__ conv_w = src.p[ly][blockIdx.y][2]; const float* __restrict__ conv_b = src.p[ly][blockIdx.y][3];
  const float* __restrict__ dt_bias = src.p[ly][blockIdx.y][4]; const float* __restrict__ A_log = src.p[ly][blockIdx.y][5];
  const float* __restrict__ Dk = src.p[ly][blockIdx.y][6]; const float* __restrict__ norm_w = src.p[ly][blockIdx.y][7];
  const int dip = 2 * di + 2 * N + H, C = di + 2 * N;
  // The two weight matrices are 99.9 % of the bytes: 32-bit index arithmetic (multiply-high division) and four
  // independent 16-byte loads in flight per thread -- the first version spent its time in 64-bit divisions
  // (0.6 TB/s on a pure copy-and-cast).
  const int d4 = d >> 2, di4 = di >> 2;
  const int n_in = dstride * d4, n_out = d * di4;                  // float4 units (checked < 2^31 on the host)
  const int nthreads = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  TW* __restrict__ win_o = Win + (long long)dir * dstride * d;
  for (int i0 = t0; i0 < n_in; i0 += 4 * nthreads) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (i < n_in && dd4.div(i) < dip) ldv<float, 4>(in_w + (long long)i * 4, v[u]);   // rows [dip, dstride) are padding
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n_in) stv<TW, 4>(win_o + (long long)i * 4, v[u]);
    }
  }
  for (int i0 = t0; i0 < n_out; i0 += 4 * nthreads) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n_out) ldv<float, 4>(out_w + (long long)i * 4, v[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n_out) {
        int r, c4; ddi4.divmod(i, r, c4);
        stv<TW, 4>(Wout + (long long)r * ndir * di + (long long)dir * di + c4 * 4, v[u]);
      }
    }
  }
  const int n_small = C * 5 + 3 * H + di;
  for (int j0 = t0; j0 < n_small; j0 += nthreads) {
    int j = j0;
    if (j < C * 4) { conv_w_o[(long long)dir * C * 4 + j] = conv_w[j]; continue; }
    j -= C * 4;
    if (j < C) { conv_b_o[(long long)dir * C + j] = conv_b[j]; continue; }
    j -= C;
    if (j < H) { dt_bias_o[dir * H + j] = dt_bias[j]; continue; }
    j -= H;
    if (j < H) { A_log_o[dir * H + j] = A_log[j]; continue; }
    j -= H;
    if (j < H) { D_o[dir * H + j] = Dk[j]; continue; }
    j -= H;
    norm_w_o[(long long)dir * di + j] = norm_w[j];
  }
}

}  // namespace hnb

using namespace hnb;

static inline int esz(int dt) { return dt == HNB_BF16 ? 2 : 4; }
static inline bool al(const void* p, int bytes) { return ((uintptr_t)p) % (size_t)bytes == 0; }

extern "C" int hnb_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, long long rows,
                                 int d, float eps, void* y, int y_dtype, float* mean, float* rstd, void* stream) {
  HNB_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows > 0 && d > 0, "layernorm_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(rows, NORM_WARPS);
  const bool v4 = d % 4 == 0 && al(x, 4 * esz(x_dtype)) && al(y, 4 * esz(y_dtype)) && al(gamma, 16) && al(beta, 16);
#define RUN(TX, TY, VN) hnb::launch_pdl(layernorm_fwd_kernel<TX, TY, VN>, dim3(grid), dim3(NORM_WARPS * 32), 0, st,  \
      (const TX*)x, gamma, beta, rows, d, eps, (TY*)y, mean, rstd)
  const int nv = cdiv(d, 128);
  if (v4 && nv <= 6) {
    const int grid2 = cdiv(rows, NORM_WARPS * 2);
#define RUNR(TX, TY, NV) hnb::launch_pdl(layernorm_fwd_reg_kernel<TX, TY, NV>, dim3(grid2), dim3(NORM_WARPS * 32), 0, st,  \
      (const TX*)x, gamma, beta, rows, d, eps, (TY*)y, mean, rstd)
    HNB_DISPATCH_DTYPE(x_dtype, TX, HNB_DISPATCH_DTYPE(y_dtype, TY, {
      if (nv <= 1) RUNR(TX, TY, 1); else if (nv <= 2) RUNR(TX, TY, 2); else if (nv <= 3) RUNR(TX, TY, 3);
      else if (nv <= 4) RUNR(TX, TY, 4); else RUNR(TX, TY, 6); }));
#undef RUNR
  } else {
    HNB_DISPATCH_DTYPE(x_dtype, TX, HNB_DISPATCH_DTYPE(y_dtype, TY, { if (v4) RUN(TX, TY, 4); else RUN(TX, TY, 1); }));
  }
#undef RUN
  HNB_LAUNCH_CHECK("layernorm_fwd");
  return HNB_OK;
}

// grid-stride kernels: exactly one wave of resident CTAs (a partial second wave would cost a full one)
template <typename K>
static int norm_grid(K kernel, long long rows, size_t smem) {
  // (device, kernel, smem) -> resident CTAs: the occupancy query is a driver call per launch otherwise
  struct E { const void* fn; size_t smem; int dev, cap; };
  static E tab[64];
  static int n = 0;
  int dev = 0, cap = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < n && !cap; ++i)
    if (tab[i].fn == (const void*)kernel && tab[i].smem == smem && tab[i].dev == dev) cap = tab[i].cap;
  if (!cap) {
    int sms = 148, occ = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NORM_WARPS * 32, smem) != cudaSuccess || occ < 1) occ = 2;
    cap = sms * occ;
    if (n < 64) tab[n++] = E{(const void*)kernel, smem, dev, cap};
  }
  const int g = cdiv(rows, NORM_WARPS);
  return g < cap ? g : cap;
}

extern "C" int hnb_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma,
                                 const float* mean, const float* rstd, const void* dres, long long rows, int d,
                                 void* dx, int dx_dtype, float* dgamma, float* dbeta, void* stream) {
  HNB_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0 && d > 0,
                "layernorm_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = d % 4 == 0 && al(x, 4 * esz(x_dtype)) && al(dy, 4 * esz(dy_dtype)) && al(dx, 4 * esz(dx_dtype)) &&
                  (!dres || al(dres, 4 * esz(dx_dtype))) && al(gamma, 16);
  const int vn = v4 ? 4 : 1;
  const int nv = cdiv(d, 32 * vn);
  HNB_CHECK_ARG(nv <= 16, "layernorm_bwd: d=%d too large", d);
  const size_t smem = (size_t)NORM_WARPS * 2 * d * sizeof(float);
#define RUN2(A, Bx, C, VN, NV) do { if (smem > 48 * 1024) HNB_CUDA_CALL(hnb_set_max_smem((const void*)layernorm_bwd_kernel<A, Bx, C, VN, NV>, (int)smem)); \
  hnb::launch_pdl(layernorm_bwd_kernel<A, Bx, C, VN, NV>, dim3(norm_grid(layernorm_bwd_kernel<A, Bx, C, VN, NV>, rows, smem)), dim3(\
      NORM_WARPS * 32), smem, st, (const A*)dy, (const Bx*)x, gamma, mean, rstd, (const C*)dres, rows, d, (C*)dx, dgamma, dbeta); } while (0)
#define RUN(A, Bx, C, VN)                                                                  \
  do {                                                                                     \
    if (nv <= 2) RUN2(A, Bx, C, VN, 2); else if (nv <= 3) RUN2(A, Bx, C, VN, 3);           \
    else if (nv <= 4) RUN2(A, Bx, C, VN, 4); else if (nv <= 6) RUN2(A, Bx, C, VN, 6);      \
    else RUN2(A, Bx, C, VN, 16);                                                           \
  } while (0)
  HNB_CHECK_ARG(x_dtype == dx_dtype, "layernorm_bwd: x and dx must share a dtype");
  HNB_DISPATCH_DTYPE(dy_dtype, TA, HNB_DISPATCH_DTYPE(x_dtype, TB, {
    if (v4) RUN(TA, TB, TB, 4); else RUN(TA, TB, TB, 1); }));
#undef RUN
#undef RUN2
  HNB_LAUNCH_CHECK("layernorm_bwd");
  return HNB_OK;
}

extern "C" int hnb_gated_norm_fwd(const void* y, const void* zxbcdt, int dtype, long long ldz, long long dstride, const int32_t* lengths,
                                  const float* norm_w, int ndir, int B, int L, int di, float eps, void* out,
                                  float* rstd, void* stream) {
  HNB_CHECK_ARG(y && zxbcdt && norm_w && out && rstd && ndir >= 1 && ndir <= 2 && B > 0 && L > 0 && di > 0,
                "gated_norm_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * L * ndir;
  const int grid = cdiv(rows, NORM_WARPS);
  const bool v4 = di % 4 == 0 && ldz % 4 == 0 && dstride % 4 == 0 && al(y, 4 * esz(dtype)) && al(zxbcdt, 4 * esz(dtype)) &&
                  al(out, 4 * esz(dtype)) && al(norm_w, 16);
  const bool v8 = dtype == HNB_BF16 && di % 8 == 0 && ldz % 8 == 0 && dstride % 8 == 0 && al(y, 16) && al(zxbcdt, 16) &&
                  al(out, 16) && al(norm_w, 16);
  const int vn = v8 ? 8 : (v4 ? 4 : 1);
  const int nv = cdiv(di, 32 * vn);
  HNB_CHECK_ARG(nv <= 16, "gated_norm_fwd: d_inner=%d too large", di);
#define RUN2(T, VN, NV) hnb::launch_pdl(gated_norm_fwd_kernel<T, VN, NV>, dim3(grid), dim3(NORM_WARPS * 32), 0, st,  \
      (const T*)y, (const T*)zxbcdt, ldz, dstride, lengths, norm_w, ndir, B, L, di, eps, (T*)out, rstd)
#define RUN(T, VN)                                                             \
  do {                                                                         \
    if (nv <= 1) RUN2(T, VN, 1); else if (nv <= 2) RUN2(T, VN, 2);             \
    else if (nv <= 3) RUN2(T, VN, 3); else if (nv <= 4) RUN2(T, VN, 4);        \
    else if (nv <= 6) RUN2(T, VN, 6); else if (nv <= 8) RUN2(T, VN, 8);        \
    else if (nv <= 12) RUN2(T, VN, 12); else RUN2(T, VN, 16);                  \
  } while (0)
  if (v8) RUN(__nv_bfloat16, 8);
  else HNB_DISPATCH_DTYPE(dtype, T, { if (v4) RUN(T, 4); else RUN(T, 1); });
#undef RUN
#undef RUN2
  HNB_LAUNCH_CHECK("gated_norm_fwd");
  return HNB_OK;
}

extern "C" int hnb_gated_norm_bwd(const void* dout, const void* y, const void* zxbcdt, int dtype, long long ldz, long long dstride,
                                  const int32_t* lengths, const float* norm_w, const float* rstd, int ndir, int B,
                                  int L, int di, void* dy, void* dzxbcdt, float* dnorm_w, void* stream) {
  HNB_CHECK_ARG(dout && y && zxbcdt && norm_w && rstd && dy && dzxbcdt && dnorm_w, "gated_norm_bwd: null pointer");
  HNB_CHECK_ARG(ndir >= 1 && ndir <= 2 && B > 0 && L > 0 && di > 0, "gated_norm_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * L;
  const bool v4 = di % 4 == 0 && ldz % 4 == 0 && dstride % 4 == 0 && al(y, 4 * esz(dtype)) && al(zxbcdt, 4 * esz(dtype)) &&
                  al(dout, 4 * esz(dtype)) && al(dy, 4 * esz(dtype)) && al(dzxbcdt, 4 * esz(dtype)) && al(norm_w, 16);
  const bool v8 = dtype == HNB_BF16 && di % 8 == 0 && ldz % 8 == 0 && dstride % 8 == 0 && al(y, 16) && al(zxbcdt, 16) &&
                  al(dout, 16) && al(dy, 16) && al(dzxbcdt, 16) && al(norm_w, 16);
  const int vn = v8 ? 8 : (v4 ? 4 : 1);
  const int nv = cdiv(di, 32 * vn);
  HNB_CHECK_ARG(nv <= 16, "gated_norm_bwd: d_inner=%d too large", di);
  const size_t smem = (size_t)NORM_WARPS * di * sizeof(float);
#define RUN2(T, VN, NV) do { if (smem > 48 * 1024) HNB_CUDA_CALL(hnb_set_max_smem((const void*)gated_norm_bwd_kernel<T, VN, NV>, (int)smem)); \
  hnb::launch_pdl(gated_norm_bwd_kernel<T, VN, NV>, dim3(dim3(std::max(1, norm_grid(gated_norm_bwd_kernel<T, VN, NV>, rows * ndir, smem) / ndir), ndir)), dim3(\
      NORM_WARPS * 32), smem, st, (const T*)dout, (const T*)y, (const T*)zxbcdt, ldz, dstride, lengths, norm_w, rstd, ndir, B, L, di, (T*)dy, (T*)dzxbcdt, dnorm_w); } while (0)
#define RUN(T, VN)                                                             \
  do {                                                                         \
    if (nv <= 1) RUN2(T, VN, 1); else if (nv <= 2) RUN2(T, VN, 2);             \
    else if (nv <= 3) RUN2(T, VN, 3); else if (nv <= 4) RUN2(T, VN, 4);        \
    else if (nv <= 6) RUN2(T, VN, 6); else if (nv <= 8) RUN2(T, VN, 8);        \
    else if (nv <= 12) RUN2(T, VN, 12); else RUN2(T, VN, 16);                  \
  } while (0)
  if (v8) RUN(__nv_bfloat16, 8);
  else HNB_DISPATCH_DTYPE(dtype, T, { if (v4) RUN(T, 4); else RUN(T, 1); });
#undef RUN
#undef RUN2
  HNB_LAUNCH_CHECK("gated_norm_bwd");
  return HNB_OK;
}

static int pack_launch(const PackSrc& src, int ndirs, int dir0, int ndir, int d, int di, int N, int H, int dstride,
                       void* Win, void* Wout, int w_dtype, float* conv_w_o, float* conv_b_o, float* dt_bias_o,
                       float* A_log_o, float* D_o, float* norm_w_o, void* stream, int nlayers = 1,
                       long long layer_stride_bytes = 0) {
  HNB_CHECK_ARG(Win && Wout && conv_w_o && conv_b_o && dt_bias_o && A_log_o && D_o && norm_w_o, "pack_mixer_params: null pointer");
  HNB_CHECK_ARG(nlayers >= 1 && nlayers <= PACK_MAX_LAYERS, "pack_mixer_params: at most %d layers per call", PACK_MAX_LAYERS);
  for (int l = 0; l < nlayers; ++l)
    for (int r = 0; r < ndirs; ++r)
      for (int k = 0; k < 8; ++k) HNB_CHECK_ARG(src.p[l][r][k] != nullptr, "pack_mixer_params: null pointer");
  HNB_CHECK_ARG(d % 4 == 0 && di % 4 == 0 && dir0 >= 0 && dir0 + ndirs <= ndir && dstride >= 2 * di + 2 * N + H,
                "pack_mixer_params: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = ((long long)dstride * d + (long long)d * di) / 4 + (long long)(di + 2 * N) * 5 + 3 * H + di;
  HNB_CHECK_ARG((long long)dstride * d < (1LL << 31) && (long long)d * di < (1LL << 31), "pack_mixer_params: weights too large");
  int gx = cdiv(total, 256 * 4);
  const int cap = 148 * 8 / (ndirs * (nlayers < 4 ? nlayers : 4));
  if (gx > cap) gx = cap > 0 ? cap : 1;
  HNB_DISPATCH_DTYPE(w_dtype, TW, (hnb::launch_pdl(pack_mixer_kernel<TW>, dim3(dim3(gx, ndirs, nlayers)), dim3(256), 0, st, src, dir0, ndir, d, di,
      FastDiv(d / 4), FastDiv(di / 4), N, H,
      dstride, (TW*)Win, (TW*)Wout, conv_w_o, conv_b_o, dt_bias_o, A_log_o, D_o, norm_w_o, layer_stride_bytes)));
  HNB_LAUNCH_CHECK("pack_mixer_params");
  return HNB_OK;
}

extern "C" int hnb_pack_mixer_params(const float* in_w, const float* out_w, const float* conv_w, const float* conv_b,
                                     const float* dt_bias, const float* A_log, const float* Dk, const float* norm_w,
                                     int dir, int ndir, int d, int di, int N, int H, int dstride, void* Win, void* Wout,
                                     int w_dtype, float* conv_w_o, float* conv_b_o, float* dt_bias_o, float* A_log_o,
                                     float* D_o, float* norm_w_o, void* stream) {
  PackSrc src = {};
  const float* one[8] = {in_w, out_w, conv_w, conv_b, dt_bias, A_log, Dk, norm_w};
  for (int k = 0; k < 8; ++k) src.p[0][0][k] = one[k];
  return pack_launch(src, 1, dir, ndir, d, di, N, H, dstride, Win, Wout, w_dtype, conv_w_o, conv_b_o, dt_bias_o, A_log_o,
                     D_o, norm_w_o, stream);
}

extern "C" int hnb_pack_mixer_params2(const float* in_w0, const float* out_w0, const float* conv_w0, const float* conv_b0,
                                      const float* dt_bias0, const float* A_log0, const float* Dk0, const float* norm_w0,
                                      const float* in_w1, const float* out_w1, const float* conv_w1, const float* conv_b1,
                                      const float* dt_bias1, const float* A_log1, const float* Dk1, const float* norm_w1,
                                      int d, int di, int N, int H, int dstride, void* Win, void* Wout, int w_dtype,
                                      float* conv_w_o, float* conv_b_o, float* dt_bias_o, float* A_log_o, float* D_o,
                                      float* norm_w_o, void* stream) {
  PackSrc src = {};
  const float* a[8] = {in_w0, out_w0, conv_w0, conv_b0, dt_bias0, A_log0, Dk0, norm_w0};
  const float* b[8] = {in_w1, out_w1, conv_w1, conv_b1, dt_bias1, A_log1, Dk1, norm_w1};
  for (int k = 0; k < 8; ++k) { src.p[0][0][k] = a[k]; src.p[0][1][k] = b[k]; }
  return pack_launch(src, 2, 0, 2, d, di, N, H, dstride, Win, Wout, w_dtype, conv_w_o, conv_b_o, dt_bias_o, A_log_o, D_o,
                     norm_w_o, stream);
}

// Every block of a stack in ONE launch (grid z = layer): 20 launches of ~18 us per encoder step were latency, not traffic
// (a step casts 248 MB of fp32 masters, ~60 us of HBM time).  params: HOST array [nlayers][ndir][8] of device pointers in the
// order of hnb_block_fwd's `params` (in_proj.w, conv1d.w, conv1d.b, dt_bias, A_log, D, norm.w, out_proj.w); packed: nlayers
// blocks of hnb_block_packed_bytes() bytes, laid out as hnb_block_fwd / hnb_block_bwd expect their `packed` argument.
extern "C" long long hnb_block_packed_layout(int d, int ndir, int di, int N, int H, int act_dtype, long long* off3);
extern "C" int hnb_pack_mixer_stack(const void* const* params, int nlayers, int ndir, int d, int di, int N, int H, int act_dtype,
                                    void* packed, void* stream) {
  HNB_CHECK_ARG(params && packed && nlayers >= 1 && (ndir == 1 || ndir == 2), "pack_mixer_stack: bad arguments");
  const int dip = 2 * di + 2 * N + H, dstride = (dip + 7) / 8 * 8, C = di + 2 * N;
  long long off[3];
  const long long per = hnb_block_packed_layout(d, ndir, di, N, H, act_dtype, off);
  uint8_t* base = static_cast<uint8_t*>(packed);
  float* sm = reinterpret_cast<float*>(base + off[2]);
  float* conv_w = sm; float* conv_b = conv_w + (size_t)ndir * C * 4; float* dt_bias = conv_b + (size_t)ndir * C;
  float* A_log = dt_bias + (size_t)ndir * H; float* Dk = A_log + (size_t)ndir * H; float* norm_w = Dk + (size_t)ndir * H;
  const float* const* P = reinterpret_cast<const float* const*>(params);
  for (int l0 = 0; l0 < nlayers; l0 += PACK_MAX_LAYERS) {
    const int nl = nlayers - l0 < PACK_MAX_LAYERS ? nlayers - l0 : PACK_MAX_LAYERS;
    PackSrc src = {};
    for (int l = 0; l < nl; ++l)
      for (int r = 0; r < ndir; ++r) {
        const float* const* q = P + ((size_t)(l0 + l) * ndir + r) * 8;
        // block order: in_w, conv_w, conv_b, dt_bias, A_log, D, norm_w, out_w  ->  kernel order: in_w, out_w, conv_w, ...
        const float* k8[8] = {q[0], q[7], q[1], q[2], q[3], q[4], q[5], q[6]};
        for (int k = 0; k < 8; ++k) src.p[l][r][k] = k8[k];
      }
    const long long shift = (long long)l0 * per;
    int rc = pack_launch(src, ndir, 0, ndir, d, di, N, H, dstride, base + shift + off[0], base + shift + off[1], act_dtype,
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(conv_w) + shift),
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(conv_b) + shift),
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dt_bias) + shift),
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(A_log) + shift),
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(Dk) + shift),
                         reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(norm_w) + shift), stream, nl, per);
    if (rc != HNB_OK) return rc;
  }
  return HNB_OK;
}
