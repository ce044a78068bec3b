// Causal depthwise conv1d (k=4) + SiLU over the xBC columns of zxbcdt, fused with dt = softplus(dt+bias)
// and with the length-aware sequence reversal of the backward-direction mixer: rows are read through
// scan_to_nat(), results are written in scan order, so no gather ever materialises a reversed copy.
//
// Bandwidth kernels: every thread owns 4 channels (one 8-byte bf16 / 16-byte fp32 vector) and a short run
// of scan positions; ALL of its row loads (run + 3 halo rows) are issued before any arithmetic so that each
// thread keeps 15-25 independent vector requests in flight (the loop-carried 4-tap window would otherwise
// serialise load -> use -> load).  Halo rows are re-read by the neighbouring tile from L2.
#include <cstdlib>

#include "common.cuh"

namespace hnb {


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
  static __device__ __forceinline__ raw_t add(const raw_t& a, const raw_t& b) {
    return make_uint4(__float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x)), __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y)),
                      __float_as_uint(__uint_as_float(a.z) + __uint_as_float(b.z)), __float_as_uint(__uint_as_float(a.w) + __uint_as_float(b.w)));
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
  static __device__ __forceinline__ raw_t add(const raw_t& a, const raw_t& b) {          // packed bf16 adds
    raw_t r;
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __hadd2(x[0], y[0]); h[1] = __hadd2(x[1], y[1]);
    return r;
  }
};

// 2 channels per thread (backward, bf16): half the taps / windows / accumulators per thread, so twice the resident warps
template <typename T> struct V8;
template <> struct V8<float> {
  static constexpr int N = 2;
  using raw_t = uint2;
  static __device__ __forceinline__ raw_t zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ raw_t ldg(const float* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ void st(float* p, const raw_t& r) { *reinterpret_cast<uint2*>(p) = r; }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) { v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); }
  static __device__ __forceinline__ raw_t pack(const float* v) { return make_uint2(__float_as_uint(v[0]), __float_as_uint(v[1])); }
  static __device__ __forceinline__ raw_t add(const raw_t& a, const raw_t& b) {
    return make_uint2(__float_as_uint(__uint_as_float(a.x) + __uint_as_float(b.x)), __float_as_uint(__uint_as_float(a.y) + __uint_as_float(b.y)));
  }
};
template <> struct V8<__nv_bfloat16> {
  static constexpr int N = 2;
  using raw_t = uint32_t;
  static __device__ __forceinline__ raw_t zero() { return 0u; }
  static __device__ __forceinline__ raw_t ldg(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const raw_t& r) { *reinterpret_cast<uint32_t*>(p) = r; }
  static __device__ __forceinline__ void unpack(const raw_t& r, float* v) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r));
    v[0] = a.x; v[1] = a.y;
  }
  static __device__ __forceinline__ raw_t pack(const float* v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ raw_t add(const raw_t& a, const raw_t& b) {
    const __nv_bfloat162 h = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};

// sigmoid: the fp32 path keeps ex2+rcp (parity 1e-3 through 20 blocks); bf16 activations take one MUFU (tanh.approx,
// abs. error ~5e-4, below the bf16 rounding of the result)
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) { return sigmoid_f(x); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// Work decomposition shared by both kernels: a block is CONV_CT column threads (4 channels each) x CONV_SEG
// time segments of one (direction, row); a thread walks its segment in tiles of TS scan positions, carrying the
// 3-row window in registers, so halo rows are read once per segment instead of once per tile.
constexpr int CONV_CT = 64;
constexpr int CONV_SEG = 4;

template <typename T, int TS, int NTILE>
__global__ void __launch_bounds__(CONV_CT * CONV_SEG)
conv_fwd_kernel(const T* __restrict__ zx, long long ldz, long long dstride, const int* __restrict__ lengths,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ dt_bias,
                int ndir, int B, int L, int di, int N, int H, T* __restrict__ xconv, float* __restrict__ dt_out) {
  pdl_enter();
  constexpr int VN = V16<T>::N;
  constexpr int RUN = TS * NTILE;
  const int ct = threadIdx.x % CONV_CT, sg = threadIdx.x / CONV_CT;
  const int dir = blockIdx.z / B, bi = blockIdx.z % B;
  const int C = di + 2 * N;
  const long long T_ = (long long)B * L;
  const int len = lengths ? min(max(lengths[bi], 0), L) : L;   // clamped like reverse_sequences (mamba_block.py:26)
  const long long xoff = (long long)dir * dstride + di;
  const long long doff = (long long)dir * dstride + di + C;
  const T* rowbase = zx + (long long)bi * L * ldz;
  const int c = (blockIdx.x * CONV_CT + ct) * VN;
  const int sb = (blockIdx.y * CONV_SEG + sg) * RUN;
  const long long sbase = (long long)dir * T_ + (long long)bi * L;

  if (c < C && sb < L) {
    float w[VN][4], bias[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4));
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = __ldg(conv_b + (long long)dir * C + c + i);
    }
    const T* src = rowbase + xoff + c;
    // Fast path (block-uniform; every segment of a fixed-length batch takes it): the segment and its halo lie inside the row
    // and on one side of `len`, so the natural index moves by a constant +-1 per scan position -- running pointers, no
    // per-position index arithmetic or bounds predicates (the kernel is issue-bound: ~40 % of its instructions were those).
    const int blk_lo = blockIdx.y * CONV_SEG * RUN - 3, blk_hi = (blockIdx.y + 1) * CONV_SEG * RUN;    // [lo, hi): the block's positions and halo
    const bool fast = blk_hi <= L && (dir == 0 || blk_hi <= len || (blk_lo >= len && blk_lo >= 0));
    if (fast) {
      const long long dstep = (dir == 1 && sb < len) ? -ldz : ldz;
      const T* pl = src + (long long)scan_to_nat(dir, sb, len) * ldz;
      typename V16<T>::raw_t h[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) h[k] = sb > 0 ? V16<T>::ldg(pl - (3 - k) * dstep) : V16<T>::zero();   // (sb = 0: warp-uniform)
      typename V16<T>::raw_t nxt[TS];
#pragma unroll
      for (int k = 0; k < TS; ++k) { nxt[k] = V16<T>::ldg(pl); pl += dstep; }
      float win[3][VN];
      V16<T>::unpack(h[0], win[0]); V16<T>::unpack(h[1], win[1]); V16<T>::unpack(h[2], win[2]);
      T* po = xconv + (sbase + sb) * C + c;
#pragma unroll
      for (int tile = 0; tile < NTILE; ++tile) {
        typename V16<T>::raw_t raw[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) raw[k] = nxt[k];
        if (tile + 1 < NTILE) {
#pragma unroll
          for (int k = 0; k < TS; ++k) { nxt[k] = V16<T>::ldg(pl); pl += dstep; }
        }
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          float cur[VN], o[VN];
          V16<T>::unpack(raw[k], cur);
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
            o[i] = pre * sigmoid_t<T>(pre);
            win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
          }
          V16<T>::st(po, V16<T>::pack(o));
          po += C;
        }
      }
    } else {
    typename V16<T>::raw_t h[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int s = sb - 3 + k;
      h[k] = s >= 0 ? V16<T>::ldg(src + (long long)scan_to_nat(dir, s, len) * ldz) : V16<T>::zero();
    }
    float win[3][VN];
    V16<T>::unpack(h[0], win[0]); V16<T>::unpack(h[1], win[1]); V16<T>::unpack(h[2], win[2]);
    // the next tile's rows are requested before the current tile's arithmetic (register double buffer)
    typename V16<T>::raw_t nxt[TS];
#pragma unroll
    for (int k = 0; k < TS; ++k)
      nxt[k] = (sb + k < L) ? V16<T>::ldg(src + (long long)scan_to_nat(dir, sb + k, len) * ldz) : V16<T>::zero();
#pragma unroll 2
    for (int tile = 0; tile < NTILE; ++tile) {
      const int s0 = sb + tile * TS;
      if (s0 >= L) break;
      typename V16<T>::raw_t raw[TS];
#pragma unroll
      for (int k = 0; k < TS; ++k) raw[k] = nxt[k];
      if (tile + 1 < NTILE) {
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          const int s = s0 + TS + k;
          nxt[k] = (s < L) ? V16<T>::ldg(src + (long long)scan_to_nat(dir, s, len) * ldz) : V16<T>::zero();
        }
      }
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        float cur[VN], o[VN];
        V16<T>::unpack(raw[k], cur);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
          o[i] = pre * sigmoid_t<T>(pre);
          win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
        }
        if (s0 + k < L) V16<T>::st(xconv + (sbase + s0 + k) * C + c, V16<T>::pack(o));
      }
    }
  }
  }
  if (blockIdx.x == 0) {                                  // dt = softplus(raw + bias): H values per row, one column block does it
    const int s_lo = blockIdx.y * CONV_SEG * RUN, s_hi = min(s_lo + CONV_SEG * RUN, L);
    for (int idx = threadIdx.x; idx < (s_hi - s_lo) * H; idx += blockDim.x) {
      const int s = s_lo + idx / H, hh = idx % H;
      const float raw = to_f(rowbase[(long long)scan_to_nat(dir, s, len) * ldz + doff + hh]);
      dt_out[(sbase + s) * H + hh] = softplus_f(raw + dt_bias[dir * H + hh]);
    }
  }
}

__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d, bool vec) {
  if (vec) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  } else {
    atomicAdd(p, a); atomicAdd(p + 1, b); atomicAdd(p + 2, c); atomicAdd(p + 3, d);
  }
}

// backward.  dout[s] for channel c comes from dxc (c < di) or dBC (c >= di), both of the activation dtype; the
// pre-activation is recomputed from zxbcdt.  d input[s'] = sum_j w[j] dpre[s'+3-j];  dw[j] = sum_s dpre[s] in[s-3+j].
// A segment owns RUN = TS*NTILE - 3 positions: its last 3 tile positions recompute the dpre halo of the next segment.
template <typename T, typename VIO, int TS, int NTILE, int MINB, bool PARTS2>
__global__ void __launch_bounds__(CONV_CT * CONV_SEG, MINB)
conv_bwd_kernel(const T* __restrict__ zx, const T* __restrict__ dxc, long long ldz, long long dstride,
                const T* __restrict__ dBC, const float* __restrict__ ddt, const int* __restrict__ lengths,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ dt_bias,
                int ndir, int B, int L, int di, int N, int H, T* __restrict__ dzx, float* __restrict__ dconv_w,
                float* __restrict__ dconv_b, float* __restrict__ ddt_bias, int vec_red, int dbc_parts,
                long long dbc_part_stride) {
  pdl_enter();
  constexpr int VN = VIO::N;
  constexpr int RUN = TS * NTILE - 3;
  static_assert(TS > 3, "layout");
  __shared__ float s_red[CONV_SEG - 1][5 * VN][CONV_CT];      // parameter-gradient partials of segments 1..3
  const int ct = threadIdx.x % CONV_CT, sg = threadIdx.x / CONV_CT;
  const int dir = blockIdx.z / B, bi = blockIdx.z % B;
  const int C = di + 2 * N;
  const long long T_ = (long long)B * L;
  const int len = lengths ? min(max(lengths[bi], 0), L) : L;   // clamped like reverse_sequences (mamba_block.py:26)
  const long long xoff = (long long)dir * dstride + di;
  const long long doff = (long long)dir * dstride + di + C;
  const T* rowbase = zx + (long long)bi * L * ldz;
  T* drowbase = dzx + (long long)bi * L * ldz;
  const long long sbase = (long long)dir * T_ + (long long)bi * L;
  const int c = (blockIdx.x * CONV_CT + ct) * VN;
  const int sb = (blockIdx.y * CONV_SEG + sg) * RUN;
  const int se = min(sb + RUN, L);

  float gw[VN][4], gb[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { gw[i][0] = gw[i][1] = gw[i][2] = gw[i][3] = 0.f; gb[i] = 0.f; }
  if (c < C && sb < L) {
    float w[VN][4], bias[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(conv_w + ((long long)dir * C + c + i) * 4));
      w[i][0] = t.x; w[i][1] = t.y; w[i][2] = t.z; w[i][3] = t.w;
      bias[i] = __ldg(conv_b + (long long)dir * C + c + i);
    }
    const bool is_x = c < di;
    const T* gsrc = is_x ? dxc + sbase * di + c : dBC + sbase * (2 * N) + (c - di);
    const long long gld = is_x ? di : 2 * N;
    const T* src = rowbase + xoff + c;
    T* dst = drowbase + xoff + c;
    // dB | dC may arrive as TWO partial sums (one per head group of the SSD dB/dC kernel).  The second part is
    // prefetched next to the first and added (packed) when the tile is consumed, a whole tile after its load was
    // issued; `two` is uniform over the block (a block's 256 channels lie on one side of the x | BC border).
    const bool two = PARTS2 && !is_x;                                // (the one-part instantiation carries no gn2)
    const T* gsrc2 = gsrc + dbc_part_stride;
    // Fast path (block-uniform; see the forward): every position this block touches lies inside the row and on one side of
    // `len` -- running pointers instead of per-position index arithmetic and bounds predicates.
    const int blk_lo = blockIdx.y * CONV_SEG * RUN - 3, blk_hi = (blockIdx.y * CONV_SEG + CONV_SEG - 1) * RUN + TS * NTILE;
    const bool fast = blk_hi <= L && (dir == 0 || blk_hi <= len || (blk_lo >= len && blk_lo >= 0));
    if (fast) {
      const long long dstep = (dir == 1 && sb < len) ? -ldz : ldz;
      const T* px = src + (long long)scan_to_nat(dir, sb, len) * ldz;
      T* pd = dst + (long long)scan_to_nat(dir, sb, len) * ldz;           // (the first store, three positions on, goes to position sb)
      const T* pg = gsrc + (long long)sb * gld;
      typename VIO::raw_t h[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) h[k] = sb > 0 ? VIO::ldg(px - (3 - k) * dstep) : VIO::zero();   // (sb = 0: warp-uniform)
      float win[3][VN], dp[3][VN];
      VIO::unpack(h[0], win[0]); VIO::unpack(h[1], win[1]); VIO::unpack(h[2], win[2]);
#pragma unroll
      for (int i = 0; i < VN; ++i) { dp[0][i] = 0.f; dp[1][i] = 0.f; dp[2][i] = 0.f; }
      typename VIO::raw_t xn[TS], gn[TS], gn2[PARTS2 ? TS : 1];
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        xn[k] = VIO::ldg(px); px += dstep;
        gn[k] = VIO::ldg(pg);
        if (PARTS2) gn2[k] = two ? VIO::ldg(pg + dbc_part_stride) : VIO::zero();
        pg += gld;
      }
#pragma unroll 2
      for (int tile = 0; tile < NTILE; ++tile) {
        const bool first = tile == 0, last = tile == NTILE - 1;
        typename VIO::raw_t xr[TS], gr[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) { xr[k] = xn[k]; gr[k] = (PARTS2 && two) ? VIO::add(gn[k], gn2[PARTS2 ? k : 0]) : gn[k]; }
        if (!last) {
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            xn[k] = VIO::ldg(px); px += dstep;
            gn[k] = VIO::ldg(pg);
            if (PARTS2 && two) gn2[k] = VIO::ldg(pg + dbc_part_stride);
            pg += gld;
          }
        }
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          float cur[VN], g[VN], dcur[VN];
          VIO::unpack(xr[k], cur);
          VIO::unpack(gr[k], g);
          const bool own = !(k >= TS - 3 && last);                        // halo positions belong to the next segment
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
            const float sgm = sigmoid_t<T>(pre);
            dcur[i] = g[i] * sgm * (1.f + pre * (1.f - sgm));
            if (own) {
              gw[i][0] += dcur[i] * win[0][i]; gw[i][1] += dcur[i] * win[1][i];
              gw[i][2] += dcur[i] * win[2][i]; gw[i][3] += dcur[i] * cur[i];
              gb[i] += dcur[i];
            }
          }
          if (!(k < 3 && first)) {                                        // d input at position s - 3 is complete
            float o[VN];
#pragma unroll
            for (int i = 0; i < VN; ++i)
              o[i] = w[i][3] * dp[0][i] + w[i][2] * dp[1][i] + w[i][1] * dp[2][i] + w[i][0] * dcur[i];
            VIO::st(pd, VIO::pack(o));
            pd += dstep;
          }
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
            dp[0][i] = dp[1][i]; dp[1][i] = dp[2][i]; dp[2][i] = dcur[i];
          }
        }
      }
    } else {
    typename VIO::raw_t h[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int s = sb - 3 + k;
      h[k] = s >= 0 ? VIO::ldg(src + (long long)scan_to_nat(dir, s, len) * ldz) : VIO::zero();
    }
    float win[3][VN], dp[3][VN];
    VIO::unpack(h[0], win[0]); VIO::unpack(h[1], win[1]); VIO::unpack(h[2], win[2]);
#pragma unroll
    for (int i = 0; i < VN; ++i) { dp[0][i] = 0.f; dp[1][i] = 0.f; dp[2][i] = 0.f; }
    typename VIO::raw_t xn[TS], gn[TS], gn2[PARTS2 ? TS : 1];                     // register double buffer (see forward)
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const bool ok = sb + k < L;
      xn[k] = ok ? VIO::ldg(src + (long long)scan_to_nat(dir, sb + k, len) * ldz) : VIO::zero();
      gn[k] = ok ? VIO::ldg(gsrc + (long long)(sb + k) * gld) : VIO::zero();
      if (PARTS2) gn2[k] = (ok && two) ? VIO::ldg(gsrc2 + (long long)(sb + k) * gld) : VIO::zero();
    }
#pragma unroll 2
    for (int tile = 0; tile < NTILE; ++tile) {
      const int s0 = sb + tile * TS;
      if (s0 >= se + 3) break;
      const bool first = tile == 0, last = tile == NTILE - 1;
      typename VIO::raw_t xr[TS], gr[TS];
#pragma unroll
      for (int k = 0; k < TS; ++k) { xr[k] = xn[k]; gr[k] = (PARTS2 && two) ? VIO::add(gn[k], gn2[PARTS2 ? k : 0]) : gn[k]; }
      if (!last) {
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          const int s = s0 + TS + k;
          const bool ok = s < L;
          xn[k] = ok ? VIO::ldg(src + (long long)scan_to_nat(dir, s, len) * ldz) : VIO::zero();
          gn[k] = ok ? VIO::ldg(gsrc + (long long)s * gld) : VIO::zero();
          if (PARTS2 && two) gn2[k] = ok ? VIO::ldg(gsrc2 + (long long)s * gld) : VIO::zero();
        }
      }
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        const int s = s0 + k;
        float cur[VN], g[VN], dcur[VN];
        VIO::unpack(xr[k], cur);
        VIO::unpack(gr[k], g);
        const bool own = !(k >= TS - 3 && last);                        // halo positions belong to the next segment
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float pre = bias[i] + w[i][0] * win[0][i] + w[i][1] * win[1][i] + w[i][2] * win[2][i] + w[i][3] * cur[i];
          const float sgm = sigmoid_t<T>(pre);
          dcur[i] = g[i] * sgm * (1.f + pre * (1.f - sgm));             // g is zero beyond L
          if (own) {
            gw[i][0] += dcur[i] * win[0][i]; gw[i][1] += dcur[i] * win[1][i];
            gw[i][2] += dcur[i] * win[2][i]; gw[i][3] += dcur[i] * cur[i];
            gb[i] += dcur[i];
          }
        }
        const int sp = s - 3;                                           // d input at sp is complete
        if (!(k < 3 && first) && sp < se) {
          float o[VN];
#pragma unroll
          for (int i = 0; i < VN; ++i)
            o[i] = w[i][3] * dp[0][i] + w[i][2] * dp[1][i] + w[i][1] * dp[2][i] + w[i][0] * dcur[i];
          VIO::st(dst + (long long)scan_to_nat(dir, sp, len) * ldz, VIO::pack(o));
        }
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          win[0][i] = win[1][i]; win[1][i] = win[2][i]; win[2][i] = cur[i];
          dp[0][i] = dp[1][i]; dp[1][i] = dp[2][i]; dp[2][i] = dcur[i];
        }
      }
    }
  }
  }
  // parameter gradients: segments 1..3 -> shared memory -> segment 0 adds and issues 5 vector reductions per thread
  if (sg > 0) {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s_red[sg - 1][i * 4 + j][ct] = gw[i][j];
      s_red[sg - 1][4 * VN + i][ct] = gb[i];
    }
  }
  __syncthreads();
  if (sg == 0 && c < C) {
#pragma unroll
    for (int q = 0; q < CONV_SEG - 1; ++q) {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gw[i][j] += s_red[q][i * 4 + j][ct];
        gb[i] += s_red[q][4 * VN + i][ct];
      }
    }
#pragma unroll
    for (int i = 0; i < VN; ++i)
      red_add4(dconv_w + ((long long)dir * C + c + i) * 4, gw[i][0], gw[i][1], gw[i][2], gw[i][3], vec_red);
    if (VN == 4) {
      red_add4(dconv_b + (long long)dir * C + c, gb[0], gb[1], gb[VN - 2], gb[VN - 1], vec_red);
    } else {
#pragma unroll
      for (int i = 0; i < VN; ++i) atomicAdd(dconv_b + (long long)dir * C + c + i, gb[i]);
    }
  }
  // dt: d raw = ddt * sigmoid(raw + bias); one column block per (row, segment group) does it.  A thread keeps ONE head (tid % H)
  // and walks positions with stride blockDim / H, so its d dt_bias contribution is a private sum; the sums meet through the
  // (by now idle) parameter-gradient scratch in shared memory.  (The first version added every element into 12 shared-memory
  // floats: 1.4 k CAS-loop atomics on 12 addresses per block, on a quarter of the blocks.)
  if (blockIdx.x == 0) {
    const int s_lo = blockIdx.y * CONV_SEG * RUN, s_hi = min(s_lo + CONV_SEG * RUN, L);
    const int G = blockDim.x / H;                                   // position lanes (H <= 64: at least 4)
    const int hh = threadIdx.x % H, g0 = threadIdx.x / H;
    float part = 0.f;
    if (g0 < G) {
      const float bias_h = dt_bias[dir * H + hh];
      for (int s = s_lo + g0; s < s_hi; s += G) {
        const long long nat = (long long)scan_to_nat(dir, s, len) * ldz + doff + hh;
        const float raw = to_f(rowbase[nat]);
        const float g = ddt[(sbase + s) * H + hh] * sigmoid_f(raw + bias_h);
        drowbase[nat] = from_f<T>(g);
        part += g;
      }
    }
    __syncthreads();                                                // s_red is no longer read by segment 0
    float* scratch = &s_red[0][0][0];                               // >= 256 floats
    scratch[threadIdx.x] = g0 < G ? part : 0.f;
    __syncthreads();
    if (threadIdx.x < H) {
      float a = 0.f;
      for (int g = 0; g < G; ++g) a += scratch[g * H + threadIdx.x];
      atomicAdd(ddt_bias + dir * H + threadIdx.x, a);
    }
  }
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

constexpr int CONV_F_TS = 4, CONV_F_NT = 4;      // forward: 16 positions per segment
constexpr int CONV_B_TS = 4, CONV_B_NT = 8;      // backward: 29 positions per segment (+3 recomputed halo)

extern "C" int hnb_conv_fwd(const void* zxbcdt, int dtype, long long ldz, long long dstride, const int32_t* lengths,
                            const float* conv_w, const float* conv_b, const float* dt_bias, int ndir, int B, int L,
                            int di, int N, int H, void* xconv, float* dt, void* stream) {
  HNB_CHECK_ARG(zxbcdt && conv_w && conv_b && dt_bias && xconv && dt, "conv_fwd: null pointer");
  int rc = conv_check("conv_fwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  const int C = di + 2 * N;
  dim3 grid(cdiv(C / 4, CONV_CT), cdiv(L, CONV_SEG * CONV_F_TS * CONV_F_NT), ndir * B);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HNB_BF16)
    hnb::launch_pdl(conv_fwd_kernel<__nv_bfloat16, CONV_F_TS, CONV_F_NT>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, 
        (const __nv_bfloat16*)zxbcdt, ldz, dstride, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H,
        (__nv_bfloat16*)xconv, dt);
  else if (dtype == HNB_F32)
    hnb::launch_pdl(conv_fwd_kernel<float, CONV_F_TS, CONV_F_NT>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, (const float*)zxbcdt, ldz, dstride,
        lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (float*)xconv, dt);
  else { set_error("conv_fwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_fwd");
  return HNB_OK;
}

extern "C" int hnb_conv_bwd(const void* zxbcdt, const void* dxc, int dtype, long long ldz, long long dstride,
                            const void* dBC, const float* ddt, const int32_t* lengths, const float* conv_w,
                            const float* conv_b, const float* dt_bias, int ndir, int B, int L, int di, int N, int H,
                            void* dzxbcdt, float* dconv_w, float* dconv_b, float* ddt_bias, int dbc_parts, void* stream) {
  HNB_CHECK_ARG(zxbcdt && dxc && dBC && ddt && conv_w && conv_b && dt_bias && dzxbcdt && dconv_w && dconv_b && ddt_bias,
                "conv_bwd: null pointer");
  HNB_CHECK_ARG(dbc_parts == 1 || (dbc_parts == 2 && dtype == HNB_BF16), "conv_bwd: dbc_parts must be 1 (or 2 with bf16 activations)");
  const long long pstride = (long long)ndir * B * L * 2 * N;
  int rc = conv_check("conv_bwd", dtype, ldz, dstride, ndir, B, L, di, N, H);
  if (rc) return rc;
  const int C = di + 2 * N;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = ((reinterpret_cast<uintptr_t>(dconv_w) | reinterpret_cast<uintptr_t>(dconv_b)) & 15) == 0;
  static const int variant = getenv("HNB_CONV_BWD_VARIANT") ? atoi(getenv("HNB_CONV_BWD_VARIANT")) : 0;   // tuning knob
  const int ysegs = cdiv(L, CONV_SEG * (CONV_B_TS * CONV_B_NT - 3));
  if (dtype == HNB_BF16 && variant == 2 && dbc_parts == 1) {
    // 2 channels per thread (64 registers, 4 resident blocks): measured SLOWER (130 vs 89 us), 4-byte accesses cost more than occupancy gains
    dim3 grid(cdiv(C / 2, CONV_CT), ysegs, ndir * B);
    hnb::launch_pdl(conv_bwd_kernel<__nv_bfloat16, V8<__nv_bfloat16>, CONV_B_TS, CONV_B_NT, 4, false>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, 
        (const __nv_bfloat16*)zxbcdt, (const __nv_bfloat16*)dxc, ldz, dstride, (const __nv_bfloat16*)dBC, ddt, lengths,
        conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (__nv_bfloat16*)dzxbcdt, dconv_w, dconv_b, ddt_bias, vec, dbc_parts, pstride);
  } else if (dtype == HNB_BF16 && dbc_parts == 2) {
    dim3 grid(cdiv(C / 4, CONV_CT), ysegs, ndir * B);
    hnb::launch_pdl(conv_bwd_kernel<__nv_bfloat16, V16<__nv_bfloat16>, CONV_B_TS, CONV_B_NT, 2, true>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, 
        (const __nv_bfloat16*)zxbcdt, (const __nv_bfloat16*)dxc, ldz, dstride, (const __nv_bfloat16*)dBC, ddt, lengths,
        conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (__nv_bfloat16*)dzxbcdt, dconv_w, dconv_b, ddt_bias, vec, dbc_parts, pstride);
  } else if (dtype == HNB_BF16) {
    dim3 grid(cdiv(C / 4, CONV_CT), ysegs, ndir * B);
    hnb::launch_pdl(conv_bwd_kernel<__nv_bfloat16, V16<__nv_bfloat16>, CONV_B_TS, CONV_B_NT, 2, false>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, 
        (const __nv_bfloat16*)zxbcdt, (const __nv_bfloat16*)dxc, ldz, dstride, (const __nv_bfloat16*)dBC, ddt, lengths,
        conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, (__nv_bfloat16*)dzxbcdt, dconv_w, dconv_b, ddt_bias, vec, dbc_parts, pstride);
  } else if (dtype == HNB_F32) {
    dim3 grid(cdiv(C / 4, CONV_CT), ysegs, ndir * B);
    hnb::launch_pdl(conv_bwd_kernel<float, V16<float>, CONV_B_TS, CONV_B_NT, 2, false>, dim3(grid), dim3(CONV_CT * CONV_SEG), 0, st, (const float*)zxbcdt,
        (const float*)dxc, ldz, dstride, (const float*)dBC, ddt, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H,
        (float*)dzxbcdt, dconv_w, dconv_b, ddt_bias, vec, dbc_parts, pstride);
  } else { set_error("conv_bwd: unsupported dtype"); return HNB_ERR_INVALID_ARG; }
  HNB_LAUNCH_CHECK("conv_bwd");
  return HNB_OK;
}
