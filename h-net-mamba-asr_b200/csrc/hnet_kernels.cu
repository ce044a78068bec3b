// H-Net dynamic-chunking stage: fused bandwidth kernels (router epilogue + ratio partials,
// single-pass boundary scan, stream compaction, EMA scan, upsample + STE) and their backwards.
// Reference semantics: /root/reference/src/dcasr/models/hnet_chunk.py (line numbers in hnet_b200.h).
#include "common.cuh"

namespace hnb {

// =============================================================================================
// router epilogue: one warp per token, q_t and k_{t-1} read once each (2*D elements per token)
// =============================================================================================
constexpr int ROUTER_WARPS = 8;

template <typename TQ, int VN>
__device__ __forceinline__ void router_dots(const TQ* __restrict__ q, const TQ* __restrict__ k, int D,
                                            int lane, float& qk, float& qq, float& kk) {
  qk = qq = kk = 0.f;
  for (int c = lane * VN; c < D; c += 32 * VN) {
    float a[VN], b[VN];
    ldv<TQ, VN>(q + c, a);
    ldv<TQ, VN>(k + c, b);
#pragma unroll
    for (int i = 0; i < VN; ++i) { qk += a[i] * b[i]; qq += a[i] * a[i]; kk += b[i] * b[i]; }
  }
  qk = warp_sum(qk); qq = warp_sum(qq); kk = warp_sum(kk);
}

template <typename TQ, typename TP, int VN>
__global__ void __launch_bounds__(ROUTER_WARPS * 32)
router_fwd_kernel(const TQ* __restrict__ qk, long long ld, const uint8_t* __restrict__ mask, int B, int L,
                  int D, float eps, TP* __restrict__ p_out, TP* __restrict__ b_out, float* __restrict__ partial) {
  pdl_enter();
  __shared__ float sred[ROUTER_WARPS][3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long n = (long long)B * L;
  const long long tok = (long long)blockIdx.x * ROUTER_WARPS + w;
  float sb = 0.f, sp = 0.f, sm = 0.f;
  if (tok < n) {
    const int t = (int)(tok % L);
    float p;
    if (t == 0) {
      p = 1.f;                                              // p_1 := 1 (hnet_chunk.py:102-103)
    } else {
      float d, qq, kk;
      router_dots<TQ, VN>(qk + tok * ld, qk + (tok - 1) * ld + D, D, lane, d, qq, kk);
      // F.cosine_similarity: normalise each (norm clamped at eps), then dot
      const float cs = d / (fmaxf(sqrtf(qq), eps) * fmaxf(sqrtf(kk), eps));
      p = fminf(fmaxf(0.5f * (1.f - cs), 0.f), 1.f);
    }
    // the decision is taken on the value as stored (bf16 routers round first, like the reference)
    const TP ps = from_f<TP>(p);
    const float pr = to_f(ps);
    float bf = (pr >= 0.5f) ? 1.f : 0.f;
    const float m = mask ? (mask[tok] ? 1.f : 0.f) : 1.f;
    if (lane == 0) {
      p_out[tok] = from_f<TP>(pr * m);
      b_out[tok] = from_f<TP>(bf * m);
    }
    sb = bf * m; sp = to_f(from_f<TP>(pr * m)) * m; sm = m;
  }
  if (lane == 0) { sred[w][0] = sb; sred[w][1] = sp; sred[w][2] = sm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < ROUTER_WARPS; ++i) { a0 += sred[i][0]; a1 += sred[i][1]; a2 += sred[i][2]; }
    partial[blockIdx.x * 4 + 0] = a0;
    partial[blockIdx.x * 4 + 1] = a1;
    partial[blockIdx.x * 4 + 2] = a2;
    partial[blockIdx.x * 4 + 3] = 0.f;
  }
}

template <typename TP>
__global__ void __launch_bounds__(ROUTER_WARPS * 32)
masked_sums_kernel(const TP* __restrict__ p, const TP* __restrict__ b, const uint8_t* __restrict__ mask, long long n,
                   float* __restrict__ partial) {
  pdl_enter();
  __shared__ float red[32];
  const long long i = (long long)blockIdx.x * (ROUTER_WARPS * 32) + threadIdx.x;
  float sb = 0.f, sp = 0.f, sm = 0.f;
  if (i < n) {
    const float m = mask ? (mask[i] ? 1.f : 0.f) : 1.f;
    sb = to_f(b[i]) * m; sp = to_f(p[i]) * m; sm = m;
  }
  sb = block_sum(sb, red); sp = block_sum(sp, red); sm = block_sum(sm, red);
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 4 + 0] = sb; partial[blockIdx.x * 4 + 1] = sp;
    partial[blockIdx.x * 4 + 2] = sm; partial[blockIdx.x * 4 + 3] = 0.f;
  }
}

// deterministic fixed-order reduction of the per-block partials + the scalar ratio loss
__global__ void ratio_finalize_kernel(const float* __restrict__ partial, int nblk, float N, float* __restrict__ stats) {
  pdl_enter();
  __shared__ double s[3][256];
  double a0 = 0, a1 = 0, a2 = 0;
  for (int i = threadIdx.x; i < nblk; i += 256) {
    a0 += partial[i * 4 + 0]; a1 += partial[i * 4 + 1]; a2 += partial[i * 4 + 2];
  }
  s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1; s[2][threadIdx.x] = a2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] += s[0][threadIdx.x + o];
      s[1][threadIdx.x] += s[1][threadIdx.x + o];
      s[2][threadIdx.x] += s[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float sumb = (float)s[0][0], sump = (float)s[1][0], summ = (float)s[2][0];
    const float den = fmaxf(summ, 1.f);
    const float F = sumb / den, G = sump / den;
    float loss = 0.f;
    if (N != 1.f) loss = (N / (N - 1.f)) * ((N - 1.f) * F * G + (1.f - F) * (1.f - G));
    stats[0] = loss; stats[1] = sumb / den; stats[2] = F; stats[3] = G; stats[4] = den;
    stats[5] = sumb; stats[6] = sump; stats[7] = summ;
  }
}

// backward: one warp per token t writes dq_t and dk_{t-1}; token 0 writes the two all-zero rows
template <typename TQ, int VN>
__global__ void __launch_bounds__(ROUTER_WARPS * 32)
router_bwd_kernel(const TQ* __restrict__ qk, long long ld, const uint8_t* __restrict__ mask, int B, int L, int D,
                  float eps, const float* __restrict__ dp_ext, const float* __restrict__ dratio,
                  const float* __restrict__ stats, float N, TQ* __restrict__ dqk) {
  pdl_enter();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long n = (long long)B * L;
  const long long tok = (long long)blockIdx.x * ROUTER_WARPS + w;
  if (tok >= n) return;
  const int t = (int)(tok % L);
  TQ* dq = dqk + tok * ld;
  if (t == 0) {                                             // p_0 is a constant; k_{L-1} is never used
    TQ* dkl = dqk + (tok + L - 1) * ld + D;
    float z[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) z[i] = 0.f;
    for (int c = lane * VN; c < D; c += 32 * VN) { stv<TQ, VN>(dq + c, z); stv<TQ, VN>(dkl + c, z); }
    return;
  }
  const TQ* q = qk + tok * ld;
  const TQ* k = qk + (tok - 1) * ld + D;
  TQ* dk = dqk + (tok - 1) * ld + D;
  float d, qq, kk;
  router_dots<TQ, VN>(q, k, D, lane, d, qq, kk);
  const float nq = sqrtf(qq), nk = sqrtf(kk);
  const float cq = fmaxf(nq, eps), ck = fmaxf(nk, eps);
  const float cs = d / (cq * ck);
  const float praw = 0.5f * (1.f - cs);
  const float m = mask ? (mask[tok] ? 1.f : 0.f) : 1.f;
  float dp = dp_ext ? dp_ext[tok] : 0.f;
  if (dratio && N != 1.f) {                                 // dL/dp = coef ((N-1)F - (1-F)) m / denom
    const float F = stats[2], den = stats[4];
    dp += dratio[0] * (N / (N - 1.f)) * ((N - 1.f) * F - (1.f - F)) * m / den;
  }
  const float pass = (praw >= 0.f && praw <= 1.f) ? 1.f : 0.f;   // clamp(0,1) gradient
  const float dcs = -0.5f * dp * m * pass;
  // cs = <q,k> / (cq ck);  d cq/dq = q/nq if nq > eps else 0 (clamp_min is not differentiated below eps)
  const float a = dcs / (cq * ck);
  const float bq = (nq > eps) ? dcs * cs / (cq * nq) : 0.f;
  const float bk = (nk > eps) ? dcs * cs / (ck * nk) : 0.f;
  for (int c = lane * VN; c < D; c += 32 * VN) {
    float qa[VN], ka[VN], o1[VN], o2[VN];
    ldv<TQ, VN>(q + c, qa);
    ldv<TQ, VN>(k + c, ka);
#pragma unroll
    for (int i = 0; i < VN; ++i) { o1[i] = a * ka[i] - bq * qa[i]; o2[i] = a * qa[i] - bk * ka[i]; }
    stv<TQ, VN>(dq + c, o1);
    stv<TQ, VN>(dk + c, o2);
  }
}

// =============================================================================================
// boundary scan: single-pass segmented inclusive scan with decoupled look-back.
// value = number of kept frames since the row start; tile state = status(2b) | flag(1b) | sum(32b)
// =============================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_MASK = 3ull << 62;
constexpr unsigned long long ST_FLAG = 1ull << 32;

struct SegPair { int sum; int flag; };
__device__ __forceinline__ SegPair seg_op(SegPair a, SegPair b) {   // a then b
  SegPair r; r.sum = b.flag ? b.sum : a.sum + b.sum; r.flag = a.flag | b.flag; return r;
}

template <typename TP>
__global__ void __launch_bounds__(SCAN_THREADS)
boundary_scan_kernel(const TP* __restrict__ bflag, long long n, int L, long long* __restrict__ memb,
                     int* __restrict__ counts, unsigned long long* tile_state, unsigned int* tile_counter) {
  pdl_enter();
  __shared__ unsigned int s_tile;
  __shared__ SegPair s_warp[SCAN_THREADS / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);       // dynamic id => forward progress
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long base = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

  int keep[SCAN_ITEMS], head[SCAN_ITEMS], val[SCAN_ITEMS];
  SegPair run{0, 0};
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const long long idx = base + i;
    keep[i] = (idx < n) ? (to_f(bflag[idx]) > 0.5f ? 1 : 0) : 0;
    head[i] = (idx < n) ? ((idx % L) == 0) : 0;
    run = seg_op(run, SegPair{keep[i], head[i]});
    val[i] = run.sum;
  }
  // warp-level segmented inclusive scan of the per-thread aggregates
  SegPair inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    SegPair up; up.sum = __shfl_up_sync(0xffffffffu, inc.sum, o); up.flag = __shfl_up_sync(0xffffffffu, inc.flag, o);
    if (lane >= o) inc = seg_op(up, inc);
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  // exclusive prefix of this thread within the block
  SegPair excl{0, 0};
  for (int i = 0; i < w; ++i) excl = seg_op(excl, s_warp[i]);
  {
    SegPair up; up.sum = __shfl_up_sync(0xffffffffu, inc.sum, 1); up.flag = __shfl_up_sync(0xffffffffu, inc.flag, 1);
    if (lane > 0) excl = seg_op(excl, up);
  }
  // tile aggregate + look-back (thread 0)
  if (threadIdx.x == 0) {
    SegPair agg{0, 0};
    for (int i = 0; i < SCAN_THREADS / 32; ++i) agg = seg_op(agg, s_warp[i]);
    volatile unsigned long long* st = tile_state;
    int prefix = 0;
    if (tile == 0) {
      st[0] = ST_PREFIX | (agg.flag ? ST_FLAG : 0ull) | (unsigned int)agg.sum;
    } else {
      st[tile] = ST_AGG | (agg.flag ? ST_FLAG : 0ull) | (unsigned int)agg.sum;
      SegPair acc{0, 0};                                            // aggregate of tiles (j, tile)
      for (int j = (int)tile - 1; j >= 0; --j) {
        unsigned long long v;
        do { v = st[j]; } while ((v & ST_MASK) == 0ull);
        SegPair tj{(int)(unsigned int)(v & 0xffffffffull), (v & ST_FLAG) ? 1 : 0};
        acc = seg_op(tj, acc);
        if ((v & ST_MASK) == ST_PREFIX || acc.flag) break;
      }
      prefix = acc.sum;
      SegPair incl = seg_op(acc, agg);
      st[tile] = ST_PREFIX | (incl.flag ? ST_FLAG : 0ull) | (unsigned int)incl.sum;
    }
    s_prefix = prefix;
  }
  __syncthreads();
  const int tile_prefix = s_prefix;
  int seen_head = excl.flag;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const long long idx = base + i;
    if (idx >= n) break;
    seen_head |= head[i];
    // carry-in: block-exclusive prefix applies until the first head in this thread's run, the tile
    // prefix applies until the first head in the tile
    int v = val[i];
    bool head_in_thread = false;
#pragma unroll
    for (int j = 0; j <= i; ++j) head_in_thread |= (head[j] != 0);
    if (!head_in_thread) v += excl.sum;
    if (!seen_head) v += tile_prefix;
    memb[idx] = (long long)max(v - 1, 0);
    if ((idx % L) == L - 1) counts[idx / L] = v;
  }
}

// =============================================================================================
// stream compaction of kept rows (one warp per frame) and its backward gather
// =============================================================================================
constexpr int ROW_WARPS = 8;

template <typename TX, typename TP, int VN>
__global__ void __launch_bounds__(ROW_WARPS * 32)
compact_rows_kernel(const TX* __restrict__ x, const TP* __restrict__ p, const TP* __restrict__ bflag,
                    const long long* __restrict__ memb, const int* __restrict__ counts, int B, int L, int D, int M,
                    TX* __restrict__ z, uint8_t* __restrict__ zmask, float* __restrict__ P, int* __restrict__ starts) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (tok >= (long long)B * L) return;
  const int bi = (int)(tok / L), t = (int)(tok % L);
  if (to_f(bflag[tok]) > 0.5f) {
    const int j = (int)memb[tok];
    const TX* src = x + tok * D;
    TX* dst = z + ((long long)bi * M + j) * D;
    for (int c = lane * VN; c < D; c += 32 * VN) {
      float r[VN]; ldv<TX, VN>(src + c, r); stv<TX, VN>(dst + c, r);
    }
    if (lane == 0) { zmask[(long long)bi * M + j] = 1; P[(long long)bi * M + j] = to_f(p[tok]); starts[(long long)bi * M + j] = t; }
  }
  if (t < M && t >= counts[bi]) {                                   // pad slot t of this row: zero fill
    TX* dst = z + ((long long)bi * M + t) * D;
    float r[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) r[i] = 0.f;
    for (int c = lane * VN; c < D; c += 32 * VN) stv<TX, VN>(dst + c, r);
    if (lane == 0) { zmask[(long long)bi * M + t] = 0; P[(long long)bi * M + t] = 0.f; starts[(long long)bi * M + t] = L; }
  }
}

template <typename TX, typename TP, int VN>
__global__ void __launch_bounds__(ROW_WARPS * 32)
compact_rows_bwd_kernel(const TX* __restrict__ dz, const TP* __restrict__ bflag, const long long* __restrict__ memb,
                        int B, int L, int D, int M, TX* __restrict__ dx, int accumulate) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (tok >= (long long)B * L) return;
  const bool keep = to_f(bflag[tok]) > 0.5f;
  if (!keep && accumulate) return;
  const int bi = (int)(tok / L);
  const TX* src = dz + ((long long)bi * M + (keep ? (int)memb[tok] : 0)) * D;
  TX* dst = dx + tok * D;
  for (int c = lane * VN; c < D; c += 32 * VN) {
    float r[VN], o[VN];
    if (keep) ldv<TX, VN>(src + c, r);
    else {
#pragma unroll
      for (int i = 0; i < VN; ++i) r[i] = 0.f;
    }
    if (accumulate) {
      ldv<TX, VN>(dst + c, o);
#pragma unroll
      for (int i = 0; i < VN; ++i) r[i] += o[i];
    }
    stv<TX, VN>(dst + c, r);
  }
}

// =============================================================================================
// EMA linear scan over the compressed sequence.  Block = (row b, tile of 32*VN channels);
// thread = (time segment, channel vector).  Each super-step covers SEGS*STEPS timesteps:
// local scan in registers -> composites through shared memory -> fix-up -> carry.
// =============================================================================================
constexpr int EMA_SEGS = 8, EMA_STEPS = 4;

// REV = false: forward recurrence out_t = a_t out_{t-1} + s_t with a_0 = 0, a_t = 1-pc_t, s_t = pc_t x_t (s_0 = x_0)
// REV = true : reverse recurrence g_t = a_{t+1} g_{t+1} + dout_t  (time index mirrored), producing dx and dP
template <typename T, bool REV, int VN>
__global__ void __launch_bounds__(EMA_SEGS * 32)
ema_kernel(const T* __restrict__ xin,      // fwd: x        bwd: dout
           const T* __restrict__ xsav,     // fwd: unused   bwd: x
           const T* __restrict__ osav,     // fwd: unused   bwd: out (forward result)
           const float* __restrict__ P, int M, int D, float pcl,
           T* __restrict__ yout,           // fwd: out      bwd: dx
           float* __restrict__ dP) {
  pdl_enter();
  __shared__ float sA[EMA_SEGS][32];
  __shared__ float sS[EMA_SEGS][32][VN];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int bi = blockIdx.y;
  const int c0 = (blockIdx.x * 32 + lane) * VN;
  const bool cvalid = c0 < D;                                       // D % VN == 0 by dispatch
  const long long rowbase = (long long)bi * M;
  float carry[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) carry[i] = 0.f;

  for (int t0 = 0; t0 < M; t0 += EMA_SEGS * EMA_STEPS) {
    float a[EMA_STEPS], loc[EMA_STEPS][VN], cumA[EMA_STEPS];
    float A = 1.f, S[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) S[i] = 0.f;
#pragma unroll
    for (int k = 0; k < EMA_STEPS; ++k) {
      const int tt = t0 + seg * EMA_STEPS + k;                      // position in scan order
      const int t = REV ? (M - 1 - tt) : tt;                        // natural index
      float av = 1.f, sv[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) sv[i] = 0.f;
      if (tt < M) {
        float xv[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) xv[i] = 0.f;
        if (cvalid) ldv<T, VN>(xin + (rowbase + t) * D + c0, xv);
        if (!REV) {
          if (t == 0) { av = 0.f;
#pragma unroll
            for (int i = 0; i < VN; ++i) sv[i] = xv[i];
          } else {
            const float pc = fminf(fmaxf(P[rowbase + t], pcl), 1.f - pcl);
            av = 1.f - pc;
#pragma unroll
            for (int i = 0; i < VN; ++i) sv[i] = pc * xv[i];
          }
        } else {
          av = (t + 1 < M) ? 1.f - fminf(fmaxf(P[rowbase + t + 1], pcl), 1.f - pcl) : 0.f;
#pragma unroll
          for (int i = 0; i < VN; ++i) sv[i] = xv[i];
        }
      }
      a[k] = av;
      A *= av;
#pragma unroll
      for (int i = 0; i < VN; ++i) { S[i] = av * S[i] + sv[i]; loc[k][i] = S[i]; }
      cumA[k] = A;
    }
    sA[seg][lane] = A;
#pragma unroll
    for (int i = 0; i < VN; ++i) sS[seg][lane][i] = S[i];
    __syncthreads();
    float inc[VN];                                                  // state entering this segment
#pragma unroll
    for (int i = 0; i < VN; ++i) inc[i] = carry[i];
    for (int s2 = 0; s2 < seg; ++s2) {
      const float As = sA[s2][lane];
#pragma unroll
      for (int i = 0; i < VN; ++i) inc[i] = As * inc[i] + sS[s2][lane][i];
    }
    // new carry = state after the last segment
    float nc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) nc[i] = inc[i];
    for (int s2 = seg; s2 < EMA_SEGS; ++s2) {
      const float As = sA[s2][lane];
#pragma unroll
      for (int i = 0; i < VN; ++i) nc[i] = As * nc[i] + sS[s2][lane][i];
    }
#pragma unroll
    for (int i = 0; i < VN; ++i) carry[i] = nc[i];
    // outputs
#pragma unroll
    for (int k = 0; k < EMA_STEPS; ++k) {
      const int tt = t0 + seg * EMA_STEPS + k;
      if (tt >= M) break;
      const int t = REV ? (M - 1 - tt) : tt;
      float o[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) o[i] = loc[k][i] + cumA[k] * inc[i];
      if (!REV) {
        if (cvalid) stv<T, VN>(yout + (rowbase + t) * D + c0, o);
      } else {
        // o = g_t.  dx_t = pc_t g_t (t>=1), dx_0 = g_0;  dpc_t = <g_t, x_t - out_{t-1}>
        float dxv[VN], part = 0.f;
        if (t == 0) {
#pragma unroll
          for (int i = 0; i < VN; ++i) dxv[i] = o[i];
        } else {
          const float Pt = P[rowbase + t];
          const float pc = fminf(fmaxf(Pt, pcl), 1.f - pcl);
          float xv[VN], ov[VN];
#pragma unroll
          for (int i = 0; i < VN; ++i) { xv[i] = 0.f; ov[i] = 0.f; }
          if (cvalid) { ldv<T, VN>(xsav + (rowbase + t) * D + c0, xv); ldv<T, VN>(osav + (rowbase + t - 1) * D + c0, ov); }
#pragma unroll
          for (int i = 0; i < VN; ++i) { dxv[i] = pc * o[i]; part += o[i] * (xv[i] - ov[i]); }
          part = warp_sum(part);
          if (lane == 0 && Pt >= pcl && Pt <= 1.f - pcl) atomicAdd(dP + rowbase + t, part);
        }
        if (cvalid) stv<T, VN>(yout + (rowbase + t) * D + c0, dxv);
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// upsample gather + confidence STE (+ residual); backward = per-chunk segment sum
// =============================================================================================
template <typename TZ, typename TY, typename TP, int VN>
__global__ void __launch_bounds__(ROW_WARPS * 32)
upsample_fwd_kernel(const TZ* __restrict__ zbar, const long long* __restrict__ memb, const TP* __restrict__ p,
                    const TP* __restrict__ bflag, const TY* __restrict__ resid, int B, int L, int D, int M,
                    TY* __restrict__ y) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (tok >= (long long)B * L) return;
  const int bi = (int)(tok / L);
  const float pv = to_f(p[tok]);
  const float c = (to_f(bflag[tok]) > 0.5f) ? pv : 1.f - pv;
  const float ste = c + (1.f - c);                                  // == 1 up to one ulp, like the reference
  const TZ* src = zbar + ((long long)bi * M + (int)memb[tok]) * D;
  for (int cc = lane * VN; cc < D; cc += 32 * VN) {
    float r[VN], o[VN];
    ldv<TZ, VN>(src + cc, r);
    if (resid) ldv<TY, VN>(resid + tok * D + cc, o);
#pragma unroll
    for (int i = 0; i < VN; ++i) r[i] = r[i] * ste + (resid ? o[i] : 0.f);
    stv<TY, VN>(y + tok * D + cc, r);
  }
}

// one warp per chunk slot (b, j): frames [start_j, start_{j+1}) (last chunk runs to L, pad frames included)
template <typename TZ, typename TY, typename TP, int VN, int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
upsample_bwd_kernel(const TY* __restrict__ dy, const TZ* __restrict__ zbar, const int* __restrict__ starts,
                    const int* __restrict__ counts, const TP* __restrict__ p, const TP* __restrict__ bflag,
                    int B, int L, int D, int M, TZ* __restrict__ dzbar, float* __restrict__ dp) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long slot = (long long)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (slot >= (long long)B * M) return;
  const int bi = (int)(slot / M), j = (int)(slot % M);
  const int nchunk = max(counts[bi], 1);
  int t0 = 0, t1 = 0;
  if (j < nchunk) {
    t0 = (j == 0) ? 0 : starts[slot];
    t1 = (j + 1 < nchunk) ? starts[slot + 1] : L;
  }
  float acc[NV][VN], zv[NV][VN];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int cc = (k * 32 + lane) * VN;
#pragma unroll
    for (int i = 0; i < VN; ++i) { acc[k][i] = 0.f; zv[k][i] = 0.f; }
    if (cc < D) ldv<TZ, VN>(zbar + slot * D + cc, zv[k]);
  }
  for (int t = t0; t < t1; ++t) {
    const long long tok = (long long)bi * L + t;
    const float pv = to_f(p[tok]);
    const bool keep = to_f(bflag[tok]) > 0.5f;
    const float c = keep ? pv : 1.f - pv;
    const float ste = c + (1.f - c);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int cc = (k * 32 + lane) * VN;
      if (cc < D) {
        float g[VN];
        ldv<TY, VN>(dy + tok * D + cc, g);
#pragma unroll
        for (int i = 0; i < VN; ++i) { acc[k][i] += g[i] * ste; dot += g[i] * zv[k][i]; }
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) dp[tok] = keep ? dot : -dot;                     // only c carries gradient: (1-c) is detached
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int cc = (k * 32 + lane) * VN;
    if (cc < D) stv<TZ, VN>(dzbar + slot * D + cc, acc[k]);
  }
}

}  // namespace hnb

// =================================================================================================
// C ABI
// =================================================================================================
using namespace hnb;

static inline bool vec_ok(const void* p, long long ld_elems, int D, int elem_bytes, int vn) {
  return (D % vn == 0) && (ld_elems % vn == 0) && (((uintptr_t)p) % (size_t)(elem_bytes * vn) == 0);
}
static inline int esize(int dt) { return dt == HNB_BF16 ? 2 : 4; }

extern "C" int hnb_router_num_partials(long long n_tokens) { return cdiv(n_tokens, ROUTER_WARPS); }

extern "C" int hnb_router_fwd(const void* qk, int qk_dtype, long long ldqk, const uint8_t* mask, int B, int L, int D,
                              float eps, void* p, void* b, int pb_dtype, float* partial, void* stream) {
  HNB_CHECK_ARG(qk && p && b && partial && B > 0 && L > 0 && D > 0 && ldqk >= 2LL * D, "router_fwd: bad arguments");
  const long long n = (long long)B * L;
  const int grid = cdiv(n, ROUTER_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(qk, ldqk, D, esize(qk_dtype), 4);
#define RUN(TQ, TP, VN) hnb::launch_pdl(router_fwd_kernel<TQ, TP, VN>, dim3(grid), dim3(ROUTER_WARPS * 32), 0, st,  \
      (const TQ*)qk, ldqk, mask, B, L, D, eps, (TP*)p, (TP*)b, partial)
  HNB_DISPATCH_DTYPE(qk_dtype, TQ, HNB_DISPATCH_DTYPE(pb_dtype, TP, { if (v4) RUN(TQ, TP, 4); else RUN(TQ, TP, 1); }));
#undef RUN
  HNB_LAUNCH_CHECK("router_fwd");
  return HNB_OK;
}

extern "C" int hnb_masked_sums(const void* p, const void* b, int pb_dtype, const uint8_t* mask, long long n,
                               float* partial, void* stream) {
  HNB_CHECK_ARG(p && b && partial && n > 0, "masked_sums: bad arguments");
  const int grid = cdiv(n, ROUTER_WARPS * 32);
  HNB_DISPATCH_DTYPE(pb_dtype, TP, (hnb::launch_pdl(masked_sums_kernel<TP>, dim3(grid), dim3(ROUTER_WARPS * 32), 0, (cudaStream_t)stream, 
      (const TP*)p, (const TP*)b, mask, n, partial)));
  HNB_LAUNCH_CHECK("masked_sums");
  return HNB_OK;
}

extern "C" int hnb_ratio_finalize(const float* partial, int nblk, float N, float* stats, void* stream) {
  HNB_CHECK_ARG(partial && stats && nblk > 0, "ratio_finalize: bad arguments");
  hnb::launch_pdl(ratio_finalize_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, partial, nblk, N, stats);
  HNB_LAUNCH_CHECK("ratio_finalize");
  return HNB_OK;
}

extern "C" int hnb_router_bwd(const void* qk, int qk_dtype, long long ldqk, const uint8_t* mask, int B, int L, int D,
                              float eps, const float* dp_ext, const float* dratio, const float* stats, float N,
                              void* dqk, void* stream) {
  HNB_CHECK_ARG(qk && dqk && B > 0 && L > 0 && D > 0 && ldqk >= 2LL * D, "router_bwd: bad arguments");
  HNB_CHECK_ARG(!dratio || stats, "router_bwd: dratio needs stats");
  const long long n = (long long)B * L;
  const int grid = cdiv(n, ROUTER_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(qk, ldqk, D, esize(qk_dtype), 4) && vec_ok(dqk, ldqk, D, esize(qk_dtype), 4);
#define RUN(TQ, VN) hnb::launch_pdl(router_bwd_kernel<TQ, VN>, dim3(grid), dim3(ROUTER_WARPS * 32), 0, st,  \
      (const TQ*)qk, ldqk, mask, B, L, D, eps, dp_ext, dratio, stats, N, (TQ*)dqk)
  HNB_DISPATCH_DTYPE(qk_dtype, TQ, { if (v4) RUN(TQ, 4); else RUN(TQ, 1); });
#undef RUN
  HNB_LAUNCH_CHECK("router_bwd");
  return HNB_OK;
}

extern "C" long long hnb_boundary_scan_ws_bytes(long long n_tokens) {
  return 8LL * (cdiv(n_tokens, SCAN_TILE) + 2);
}

extern "C" int hnb_boundary_scan(const void* b, int pb_dtype, int B, int L, int64_t* membership, int32_t* counts,
                                 void* ws, void* stream) {
  HNB_CHECK_ARG(b && membership && counts && ws && B > 0 && L > 0, "boundary_scan: bad arguments");
  const long long n = (long long)B * L;
  const int tiles = cdiv(n, SCAN_TILE);
  cudaStream_t st = (cudaStream_t)stream;
  HNB_CUDA_CALL(cudaMemsetAsync(ws, 0, (size_t)hnb_boundary_scan_ws_bytes(n), st));
  unsigned long long* state = (unsigned long long*)ws + 1;
  unsigned int* counter = (unsigned int*)ws;
  HNB_DISPATCH_DTYPE(pb_dtype, TP, (hnb::launch_pdl(boundary_scan_kernel<TP>, dim3(tiles), dim3(SCAN_THREADS), 0, st, 
      (const TP*)b, n, L, (long long*)membership, counts, state, counter)));
  HNB_LAUNCH_CHECK("boundary_scan");
  return HNB_OK;
}

extern "C" int hnb_compact_rows(const void* x, int x_dtype, const void* p, const void* b, int pb_dtype,
                                const int64_t* membership, const int32_t* counts, int B, int L, int D, int M,
                                void* z, uint8_t* z_mask, float* P, int32_t* starts, void* stream) {
  HNB_CHECK_ARG(x && p && b && membership && counts && z && z_mask && P && starts, "compact_rows: null pointer");
  HNB_CHECK_ARG(B > 0 && L > 0 && D > 0 && M >= 1 && M <= L, "compact_rows: bad sizes (need 1 <= M <= L)");
  const int grid = cdiv((long long)B * L, ROW_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(x, D, D, esize(x_dtype), 4) && vec_ok(z, D, D, esize(x_dtype), 4);
#define RUN(TX, TP, VN) hnb::launch_pdl(compact_rows_kernel<TX, TP, VN>, dim3(grid), dim3(ROW_WARPS * 32), 0, st,  \
      (const TX*)x, (const TP*)p, (const TP*)b, (const long long*)membership, counts, B, L, D, M, (TX*)z, z_mask, P, starts)
  HNB_DISPATCH_DTYPE(x_dtype, TX, HNB_DISPATCH_DTYPE(pb_dtype, TP, { if (v4) RUN(TX, TP, 4); else RUN(TX, TP, 1); }));
#undef RUN
  HNB_LAUNCH_CHECK("compact_rows");
  return HNB_OK;
}

extern "C" int hnb_compact_rows_bwd(const void* dz, int dtype, const void* b, int pb_dtype, const int64_t* membership,
                                    int B, int L, int D, int M, void* dx, int accumulate, void* stream) {
  HNB_CHECK_ARG(dz && b && membership && dx && B > 0 && L > 0 && D > 0 && M >= 1, "compact_rows_bwd: bad arguments");
  const int grid = cdiv((long long)B * L, ROW_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(dz, D, D, esize(dtype), 4) && vec_ok(dx, D, D, esize(dtype), 4);
#define RUN(TX, TP, VN) hnb::launch_pdl(compact_rows_bwd_kernel<TX, TP, VN>, dim3(grid), dim3(ROW_WARPS * 32), 0, st,  \
      (const TX*)dz, (const TP*)b, (const long long*)membership, B, L, D, M, (TX*)dx, accumulate)
  HNB_DISPATCH_DTYPE(dtype, TX, HNB_DISPATCH_DTYPE(pb_dtype, TP, { if (v4) RUN(TX, TP, 4); else RUN(TX, TP, 1); }));
#undef RUN
  HNB_LAUNCH_CHECK("compact_rows_bwd");
  return HNB_OK;
}

extern "C" int hnb_ema_fwd(const void* x, int dtype, const float* P, int B, int M, int D, float p_clamp, void* out,
                           void* stream) {
  HNB_CHECK_ARG(x && P && out && B > 0 && M > 0 && D > 0, "ema_fwd: bad arguments");
  const int vn = vec_ok(x, D, D, esize(dtype), 2) && vec_ok(out, D, D, esize(dtype), 2) ? 2 : 1;
  dim3 grid(cdiv(D, 32 * vn), B);
  cudaStream_t st = (cudaStream_t)stream;
#define RUN(T, VN) hnb::launch_pdl(ema_kernel<T, false, VN>, dim3(grid), dim3(EMA_SEGS * 32), 0, st,  \
      (const T*)x, nullptr, nullptr, P, M, D, p_clamp, (T*)out, nullptr)
  HNB_DISPATCH_DTYPE(dtype, T, { if (vn == 2) RUN(T, 2); else RUN(T, 1); });
#undef RUN
  HNB_LAUNCH_CHECK("ema_fwd");
  return HNB_OK;
}

extern "C" int hnb_ema_bwd(const void* dout, const void* x, const void* out, int dtype, const float* P, int B, int M,
                           int D, float p_clamp, void* dx, float* dP, void* stream) {
  HNB_CHECK_ARG(dout && x && out && P && dx && dP && B > 0 && M > 0 && D > 0, "ema_bwd: bad arguments");
  const int vn = vec_ok(x, D, D, esize(dtype), 2) && vec_ok(out, D, D, esize(dtype), 2) &&
                 vec_ok(dout, D, D, esize(dtype), 2) && vec_ok(dx, D, D, esize(dtype), 2) ? 2 : 1;
  dim3 grid(cdiv(D, 32 * vn), B);
  cudaStream_t st = (cudaStream_t)stream;
#define RUN(T, VN) hnb::launch_pdl(ema_kernel<T, true, VN>, dim3(grid), dim3(EMA_SEGS * 32), 0, st,  \
      (const T*)dout, (const T*)x, (const T*)out, P, M, D, p_clamp, (T*)dx, dP)
  HNB_DISPATCH_DTYPE(dtype, T, { if (vn == 2) RUN(T, 2); else RUN(T, 1); });
#undef RUN
  HNB_LAUNCH_CHECK("ema_bwd");
  return HNB_OK;
}

extern "C" int hnb_upsample_fwd(const void* zbar, int z_dtype, const int64_t* membership, const void* p, const void* b,
                                int pb_dtype, const void* resid, int B, int L, int D, int M, void* y, int y_dtype,
                                void* stream) {
  HNB_CHECK_ARG(zbar && membership && p && b && y && B > 0 && L > 0 && D > 0 && M >= 1, "upsample_fwd: bad arguments");
  const int grid = cdiv((long long)B * L, ROW_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(zbar, D, D, esize(z_dtype), 4) && vec_ok(y, D, D, esize(y_dtype), 4) &&
                  (!resid || vec_ok(resid, D, D, esize(y_dtype), 4));
#define RUN(TZ, TY, TP, VN) hnb::launch_pdl(upsample_fwd_kernel<TZ, TY, TP, VN>, dim3(grid), dim3(ROW_WARPS * 32), 0, st,  \
      (const TZ*)zbar, (const long long*)membership, (const TP*)p, (const TP*)b, (const TY*)resid, B, L, D, M, (TY*)y)
  HNB_DISPATCH_DTYPE(z_dtype, TZ, HNB_DISPATCH_DTYPE(y_dtype, TY, HNB_DISPATCH_DTYPE(pb_dtype, TP, {
    if (v4) RUN(TZ, TY, TP, 4); else RUN(TZ, TY, TP, 1); })));
#undef RUN
  HNB_LAUNCH_CHECK("upsample_fwd");
  return HNB_OK;
}

extern "C" int hnb_upsample_bwd(const void* dy, int y_dtype, const void* zbar, int z_dtype, const int64_t* membership,
                                const int32_t* starts, const int32_t* counts, const void* p, const void* b,
                                int pb_dtype, int B, int L, int D, int M, void* dzbar, float* dp, void* stream) {
  (void)membership;
  HNB_CHECK_ARG(dy && zbar && starts && counts && p && b && dzbar && dp, "upsample_bwd: null pointer");
  HNB_CHECK_ARG(B > 0 && L > 0 && D > 0 && M >= 1, "upsample_bwd: bad sizes");
  const int grid = cdiv((long long)B * M, ROW_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = vec_ok(zbar, D, D, esize(z_dtype), 4) && vec_ok(dy, D, D, esize(y_dtype), 4) &&
                  vec_ok(dzbar, D, D, esize(z_dtype), 4);
  const int vn = v4 ? 4 : 1;
  const int nv = cdiv(D, 32 * vn);
  HNB_CHECK_ARG(nv <= 16, "upsample_bwd: D=%d too large for the register tile", D);
#define RUN2(TZ, TY, TP, VN, NV) hnb::launch_pdl(upsample_bwd_kernel<TZ, TY, TP, VN, NV>, dim3(grid), dim3(ROW_WARPS * 32), 0, st,  \
      (const TY*)dy, (const TZ*)zbar, starts, counts, (const TP*)p, (const TP*)b, B, L, D, M, (TZ*)dzbar, dp)
#define RUN(TZ, TY, TP, VN)                                                 \
  do {                                                                      \
    if (nv <= 1) RUN2(TZ, TY, TP, VN, 1); else if (nv <= 2) RUN2(TZ, TY, TP, VN, 2);       \
    else if (nv <= 4) RUN2(TZ, TY, TP, VN, 4); else if (nv <= 8) RUN2(TZ, TY, TP, VN, 8);  \
    else RUN2(TZ, TY, TP, VN, 16);                                          \
  } while (0)
  HNB_DISPATCH_DTYPE(z_dtype, TZ, HNB_DISPATCH_DTYPE(y_dtype, TY, HNB_DISPATCH_DTYPE(pb_dtype, TP, {
    if (v4) RUN(TZ, TY, TP, 4); else RUN(TZ, TY, TP, 1); })));
#undef RUN
#undef RUN2
  HNB_LAUNCH_CHECK("upsample_bwd");
  return HNB_OK;
}
