// SSD (Mamba-2 selective scan) in scan order, chunked dual form, fp32 arithmetic on CUDA cores.
// This is the exact path (impl 0): it serves fp32 activations (decode / parity) and is the numerical
// yard-stick for the tcgen05 kernels (impl 1, ssd_tcgen05.cu).
//
//   h_t = exp(dt_t A) h_{t-1} + dt_t B_t (x) x_t ,  y_t = C_t . h_t + D x_t      (per head; B, C shared)
//
// Chunk Q = 64.  Forward:  chunk_state -> state_pass -> chunk_scan.
// Backward: chunk_state(dY, C) -> state_pass(reverse) -> bwd_chunk.
// Every small matmul is written as an outer-product loop over the contraction index k with the
// lane-varying operand stored [k][lanes] in shared memory (conflict-free) and the other operand
// broadcast.  Per-chunk states are stored [N][P] (p contiguous) so that no state tile is ever transposed.
#include <cstdlib>

#include "common.cuh"

namespace hnb {

constexpr int SQ = 64;      // chunk length
constexpr int SP = 64;      // head dim
constexpr int SN = 128;     // state dim
constexpr int ST = 256;     // threads per CTA
constexpr int PADQ = SQ + 1;

// cumulative log-decay of one (row, head, chunk): s_dt[q] = dt_q (0 beyond L), s_cs[q] = inclusive cumsum(dt*A)
__device__ __forceinline__ void chunk_cumsum(const float* __restrict__ dtp, long long stride, int q_valid, float A,
                                             float* s_dt, float* s_cs) {
  const int tid = threadIdx.x;
  if (tid < SQ) {
    const float d = (tid < q_valid) ? dtp[(long long)tid * stride] : 0.f;
    s_dt[tid] = d;
    float v = d * A;
    const int lane = tid & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    s_cs[tid] = v;
  }
  __syncthreads();
  if (tid >= 32 && tid < SQ) s_cs[tid] += s_cs[31];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// chunk_state: out[n][p] = sum_q w_q V[q][n] U[q][p]
//   MODE 0 (forward):  U = x,  V = B,  w_q = exp(cs_last - cs_q) dt_q ; also writes decay = exp(cs_last)
//   MODE 1 (backward): U = dy, V = C,  w_q = exp(cs_q)
// grid (nchunks, ndir*B), loop over heads.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(ST, 4)                        // 64 registers, 48.5 KB: four CTAs per SM
ssd_chunk_state_kernel(const T* __restrict__ U, long long ldu, const T* __restrict__ xconv, int C, int di,
                       const float* __restrict__ dt, const float* __restrict__ A_log, int ndir, int B, int L, int H,
                       int nc, int hpg, float* __restrict__ states, float* __restrict__ decay) {
  extern __shared__ float smem[];
  float* Vs = smem;                       // [SQ][SN]
  float* Us = Vs + SQ * SN;               // [SQ][SP] (pre-scaled by w_q)
  float* s_dt = Us + SQ * SP;             // [SQ]
  float* s_cs = s_dt + SQ;                // [SQ]
  const int c = blockIdx.z, db = blockIdx.y, dir = db / B;      // grid (head groups, ndir*B, chunks): the short last chunks run last
  const int h_lo = blockIdx.x * hpg, h_hi = min(H, h_lo + hpg);
  const int tid = threadIdx.x;
  const int q0 = c * SQ, qv = min(SQ, L - q0);
  const long long row0 = (long long)db * L + q0;
  const int voff = di + (MODE == 0 ? 0 : SN);
  for (int i = tid; i < SQ * SN / 4; i += ST) {
    const int q = i / (SN / 4), n4 = (i % (SN / 4)) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < qv) ldv<T, 4>(xconv + (row0 + q) * C + voff + n4, v);
    *reinterpret_cast<float4*>(Vs + q * SN + n4) = make_float4(v[0], v[1], v[2], v[3]);
  }
  // thread tile: 8 consecutive state rows n x 4 consecutive columns p.  Both operands of the outer-product step are k-major in
  // shared memory, so a step is three 128-bit loads (two of them warp-wide broadcasts) for 32 FMAs; the first version
  // (strided 8 x 4 tile, twelve 32-bit loads per step) was bound by the shared-memory pipe, not by the FMA pipe.
  const int tp = tid & 15, tn = tid >> 4;
  const float4* Va = reinterpret_cast<const float4*>(Vs) + 2 * tn;
  const float4* Ub = reinterpret_cast<const float4*>(Us) + tp;
  for (int h = h_lo; h < h_hi; ++h) {
    const float A = -__expf(A_log[dir * H + h]);
    __syncthreads();
    chunk_cumsum(dt + row0 * H + h, H, qv, A, s_dt, s_cs);
    const float cs_last = s_cs[SQ - 1];
    for (int i = tid; i < SQ * SP / 4; i += ST) {
      const int q = i / (SP / 4), p4 = (i % (SP / 4)) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < qv) ldv<T, 4>(U + (row0 + q) * ldu + h * SP + p4, v);
      const float w = (MODE == 0) ? __expf(cs_last - s_cs[q]) * s_dt[q] : __expf(s_cs[q]);
      *reinterpret_cast<float4*>(Us + q * SP + p4) = make_float4(v[0] * w, v[1] * w, v[2] * w, v[3] * w);
    }
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int q = 0; q < qv; ++q) {                                    // frames beyond the row's end contribute nothing
      const float4 a0 = Va[q * (SN / 4)], a1 = Va[q * (SN / 4) + 1], bv = Ub[q * (SP / 4)];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    float* out = states + (((long long)db * nc + c) * H + h) * (SN * SP);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(out + (8 * tn + i) * SP + 4 * tp) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (MODE == 0 && tid == 0) decay[((long long)db * H + h) * nc + c] = __expf(cs_last);
  }
}

// ---------------------------------------------------------------------------------------------
// inter-chunk state pass (elementwise over the 128x64 state, sequential over chunks).
//   forward: states[c] <- state entering chunk c;           running = decay_c running + local_c
//   reverse: states[c] <- gradient w.r.t. the state LEAVING chunk c; running = decay_c running + local_c
// ---------------------------------------------------------------------------------------------
template <bool REV>
__global__ void __launch_bounds__(256)
ssd_state_pass_kernel(float* __restrict__ states, const float* __restrict__ decay, int H, int nc, long long total4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;      // float4 index over [db][h][N*P/4]
  if (i >= total4) return;
  constexpr int PER = SN * SP / 4;
  const long long dbh = i / PER;
  const int e = (int)(i % PER);
  const long long db = dbh / H;
  const int h = (int)(dbh % H);
  float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
  // chunks in batches of 8: every load of a batch is issued before its first store (the in-place update otherwise makes each
  // chunk a dependent DRAM round trip: 1.2 TB/s measured)
  constexpr int BT = 8;
  float4* base = reinterpret_cast<float4*>(states + ((long long)db * nc * H + h) * (SN * SP)) + e;
  const long long cstride = (long long)H * (SN * SP / 4);
  const float* dk = decay + ((long long)db * H + h) * nc;
  for (int k0 = 0; k0 < nc; k0 += BT) {
    float4 loc[BT];
    float d[BT];
#pragma unroll
    for (int u = 0; u < BT; ++u) {
      const int k = k0 + u;
      if (k < nc) {
        const int c = REV ? (nc - 1 - k) : k;
        loc[u] = __ldcs(base + c * cstride);
        d[u] = __ldg(dk + c);
      }
    }
#pragma unroll
    for (int u = 0; u < BT; ++u) {
      const int k = k0 + u;
      if (k < nc) {
        const int c = REV ? (nc - 1 - k) : k;
        base[c * cstride] = run;
        run.x = d[u] * run.x + loc[u].x; run.y = d[u] * run.y + loc[u].y;
        run.z = d[u] * run.z + loc[u].z; run.w = d[u] * run.w + loc[u].w;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// chunk_scan (forward output): y = (L o C B^T)(dt x) + exp(cs) C S_in + D x
// grid (nchunks, ndir*B), loop over heads; G = C B^T is computed once and kept in registers.
// ---------------------------------------------------------------------------------------------
constexpr int MTS = SQ + 4;                                           // row stride of the transposed score tile (2-way store conflicts)
constexpr int SCAN_SMEM_FLOATS = SN * SQ + (SN * PADQ > SQ * MTS + SQ * SP + SN * SP ? SN * PADQ : SQ * MTS + SQ * SP + SN * SP) + 2 * SQ;

// Thread tile: 4 consecutive rows t x 4 consecutive columns p.  Every operand of the three products is stored k-major
// (C and the score tile transposed: Ct[n][t], Mt[s][t]), so an outer-product step is two 128-bit shared-memory loads for 16
// FMAs (the first version: eight 32-bit loads, bound by the shared-memory pipe at one CTA per SM).  B^T is only needed for
// G = C B^T at the start and shares its memory with the per-head tiles: 98 KB, two CTAs per SM.  Rows beyond the row's end and
// the all-zero upper triangle of the score tile are skipped.
template <typename T>
__global__ void __launch_bounds__(ST, 2)
ssd_chunk_scan_kernel(const T* __restrict__ xconv, int C, int di, const float* __restrict__ dt,
                      const float* __restrict__ A_log, const float* __restrict__ Dskip,
                      const float* __restrict__ states, int ndir, int B, int L, int H, int nc, int hpg,
                      T* __restrict__ y) {
  extern __shared__ float smem[];
  float* Ct = smem;                        // [SN][SQ]   C transposed
  float* Bt = Ct + SN * SQ;                // [SN][PADQ] B transposed (dead after G)
  float* Mt = Bt;                          // [SQ][MTS]  score tile transposed: Mt[s][t]
  float* Xs = Mt + SQ * MTS;               // [SQ][SP]
  float* Ss = Xs + SQ * SP;                // [SN][SP]
  float* s_dt = smem + SCAN_SMEM_FLOATS - 2 * SQ;
  float* s_cs = s_dt + SQ;
  const int c = blockIdx.z, db = blockIdx.y, dir = db / B;      // grid (head groups, ndir*B, chunks): the short last chunks run last
  const int h_lo = blockIdx.x * hpg, h_hi = min(H, h_lo + hpg);
  const int tid = threadIdx.x;
  const int q0 = c * SQ, qv = min(SQ, L - q0);
  const long long row0 = (long long)db * L + q0;
  for (int i = tid; i < SQ * SN / 4; i += ST) {                       // q fastest: conflict-free transposed stores
    const int q = i % SQ, n4 = (i / SQ) * 4;
    float vb[4] = {0.f, 0.f, 0.f, 0.f}, vc[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < qv) { ldv<T, 4>(xconv + (row0 + q) * C + di + n4, vb); ldv<T, 4>(xconv + (row0 + q) * C + di + SN + n4, vc); }
#pragma unroll
    for (int k = 0; k < 4; ++k) { Ct[(n4 + k) * SQ + q] = vc[k]; Bt[(n4 + k) * PADQ + q] = vb[k]; }
  }
  __syncthreads();
  const int tj = tid & 15, ti = tid >> 4;            // rows t = 4 ti + i; G columns q = tj + 16 j; output columns p = 4 tj + j
  const bool live = 4 * ti < qv;                     // this thread owns at least one valid row
  const float4* Ca = reinterpret_cast<const float4*>(Ct) + ti;
  float G[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) G[i][j] = 0.f;
  if (live) {
#pragma unroll 4
    for (int n = 0; n < SN; ++n) {
      const float4 av = Ca[n * (SQ / 4)];
      const float a[4] = {av.x, av.y, av.z, av.w};
      float b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bt[n * PADQ + tj + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[i][j] = fmaf(a[i], b[j], G[i][j]);
    }
  }
  const float4* Ma = reinterpret_cast<const float4*>(Mt) + ti;
  const float4* Xb = reinterpret_cast<const float4*>(Xs) + tj;
  const float4* Sb = reinterpret_cast<const float4*>(Ss) + tj;
  const int s_end = min(qv, 4 * ti + 4);             // M[t][s] = 0 for s > t
  for (int h = h_lo; h < h_hi; ++h) {
    const float A = -__expf(A_log[dir * H + h]);
    const float Dh = Dskip[dir * H + h];
    __syncthreads();                                 // B^T (first head) / the previous head's tiles are no longer read
    chunk_cumsum(dt + row0 * H + h, H, qv, A, s_dt, s_cs);
    for (int i = tid; i < SQ * SP / 4; i += ST) {
      const int q = i / (SP / 4), p4 = (i % (SP / 4)) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < qv) ldv<T, 4>(xconv + (row0 + q) * C + h * SP + p4, v);
      *reinterpret_cast<float4*>(Xs + q * SP + p4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    const float4* sg = reinterpret_cast<const float4*>(states + (((long long)db * nc + c) * H + h) * (SN * SP));
    for (int i = tid; i < SN * SP / 4; i += ST) reinterpret_cast<float4*>(Ss)[i] = __ldcs(sg + i);
    {
      float cst[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) cst[i] = s_cs[4 * ti + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int sc = tj + 16 * j;
        const float css = s_cs[sc], dts = s_dt[sc];
        float m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = (sc <= 4 * ti + i) ? G[i][j] * __expf(cst[i] - css) * dts : 0.f;
        *reinterpret_cast<float4*>(Mt + sc * MTS + 4 * ti) = make_float4(m[0], m[1], m[2], m[3]);
      }
    }
    __syncthreads();
    if (live) {
      float acc[4][4], off[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; off[i][j] = 0.f; }
#pragma unroll 4
      for (int sq = 0; sq < s_end; ++sq) {
        const float4 av = Ma[sq * (MTS / 4)], bv = Xb[sq * (SP / 4)];
        const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
#pragma unroll 4
      for (int n = 0; n < SN; ++n) {
        const float4 av = Ca[n * (SQ / 4)], bv = Sb[n * (SP / 4)];
        const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) off[i][j] = fmaf(a[i], b[j], off[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = 4 * ti + i;
        if (t < qv) {
          const float e = __expf(s_cs[t]);
          const float4 xv = Xb[t * (SP / 4)];
          const float o[4] = {acc[i][0] + e * off[i][0] + Dh * xv.x, acc[i][1] + e * off[i][1] + Dh * xv.y,
                              acc[i][2] + e * off[i][2] + Dh * xv.z, acc[i][3] + e * off[i][3] + Dh * xv.w};
          stv<T, 4>(y + (row0 + t) * di + h * SP + 4 * tj, o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward chunk kernel.  grid (nchunks, ndir*B), loop over heads.
// Per head (x, dy, y of the head; S_in = state entering the chunk, Gst = d loss / d state leaving it):
//   K[t,q] = G[t,q] e^{cs_t-cs_q}  (t>=q);   W[t,q] = <dy_t, x_q> dt_q e^{cs_t-cs_q}  (t>=q)
//   du[q,p] = sum_t K[t,q] dy[t,p] + e^{cs_last-cs_q} sum_n B[q,n] Gst[n,p]
//   dx = dt du + D dy ;  ddt[q] = <du_q, x_q> + A * sum_{t>=q} d cs_t, with d cs gathered term by term
//   (score matrix, Y_off, chunk state, inter-chunk decay) -- no difference of long sums, so bf16 inputs are safe
//   dC[t,n] += sum_q W[t,q] B[q,n] + e^{cs_t} sum_p dy[t,p] S_in[n,p]
//   dB[q,n] += sum_t W[t,q] C[t,n] + e^{cs_last-cs_q} dt_q sum_p x[q,p] Gst[n,p]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ST)
ssd_bwd_chunk_kernel(const T* __restrict__ dy, const T* __restrict__ xconv, int C, int di,
                     const float* __restrict__ dt, const float* __restrict__ A_log, const float* __restrict__ Dskip,
                     const float* __restrict__ states, const float* __restrict__ gstates, int ndir, int B, int L,
                     int H, int nc, T* __restrict__ dxc, T* __restrict__ dBC, float* __restrict__ ddt,
                     float* __restrict__ dA_log, float* __restrict__ dD) {
  extern __shared__ float smem[];
  float* Bn = smem;                        // [SQ][SN]
  float* Cn = Bn + SQ * SN;                // [SQ][SN]
  float* Al = Cn + SQ * SN;                // aliased: Bt [SN][PADQ] first, then dYt [SP][PADQ] + Xt [SP][PADQ]
  float* Bt = Al;
  float* dYt = Al;
  float* Xt = Al + SP * PADQ;
  float* Gs = Al + SN * PADQ;              // [SQ][SQ]  C B^T
  float* dYs = Gs + SQ * SQ;               // [SQ][SP]
  float* Ks = dYs + SQ * SP;               // [SQ][SQ]
  float* Ws = Ks + SQ * SQ;                // [SQ][SQ]
  float* s_dt = Ws + SQ * SQ;
  float* s_cs = s_dt + SQ;
  float* s_red = s_cs + SQ;                // [32]
  float* s_dcs = s_red + 32;               // [SQ]  d loss / d cs_t  (cs = inclusive cumulative log decay)
  float* s_ddtx = s_dcs + SQ;              // [SQ]  <du_q, x_q>
  const int c = blockIdx.x, db = blockIdx.y, dir = db / B;
  const int tid = threadIdx.x;
  const int q0 = c * SQ, qv = min(SQ, L - q0);
  const long long row0 = (long long)db * L + q0;
  for (int i = tid; i < SQ * SN / 4; i += ST) {
    const int q = i / (SN / 4), n4 = (i % (SN / 4)) * 4;
    float vb[4] = {0.f, 0.f, 0.f, 0.f}, vc[4] = {0.f, 0.f, 0.f, 0.f};
    if (q < qv) { ldv<T, 4>(xconv + (row0 + q) * C + di + n4, vb); ldv<T, 4>(xconv + (row0 + q) * C + di + SN + n4, vc); }
    *reinterpret_cast<float4*>(Bn + q * SN + n4) = make_float4(vb[0], vb[1], vb[2], vb[3]);
    *reinterpret_cast<float4*>(Cn + q * SN + n4) = make_float4(vc[0], vc[1], vc[2], vc[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) Bt[(n4 + k) * PADQ + q] = vb[k];
  }
  __syncthreads();
  const int tj = tid & 15, ti = tid >> 4;           // 64x64 tiles
  const int wn = tid & 31, wt = tid >> 5;           // 64x128 tiles: rows wt+8i, cols wn+32j
  {
    float g[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
    for (int n = 0; n < SN; ++n) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Cn[(ti + 16 * i) * SN + n];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bt[n * PADQ + tj + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) g[i][j] += a[i] * b[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Gs[(ti + 16 * i) * SQ + tj + 16 * j] = g[i][j];
  }
  float accC[8][4], accB[8][4], accC2[8][4], accB2[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { accC[i][j] = 0.f; accB[i][j] = 0.f; accC2[i][j] = 0.f; accB2[i][j] = 0.f; }

  for (int h = 0; h < H; ++h) {
    const float A = -__expf(A_log[dir * H + h]);
    const float Dh = Dskip[dir * H + h];
    __syncthreads();                                // everyone is done with Bt / the previous head's tiles
    if (tid < SQ) { s_dcs[tid] = 0.f; s_ddtx[tid] = 0.f; }
    chunk_cumsum(dt + row0 * H + h, H, qv, A, s_dt, s_cs);
    const float cs_last = s_cs[SQ - 1];
    for (int i = tid; i < SQ * SP / 4; i += ST) {
      const int q = i / (SP / 4), p4 = (i % (SP / 4)) * 4;
      float vx[4] = {0.f, 0.f, 0.f, 0.f}, vd[4] = {0.f, 0.f, 0.f, 0.f};
      if (q < qv) { ldv<T, 4>(xconv + (row0 + q) * C + h * SP + p4, vx); ldv<T, 4>(dy + (row0 + q) * di + h * SP + p4, vd); }
      *reinterpret_cast<float4*>(dYs + q * SP + p4) = make_float4(vd[0], vd[1], vd[2], vd[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) { Xt[(p4 + k) * PADQ + q] = vx[k]; dYt[(p4 + k) * PADQ + q] = vd[k]; }
    }
    __syncthreads();
    // K and W (64x64, rows t, cols q)
    {
      float r[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r[i][j] = 0.f;
      for (int p = 0; p < SP; ++p) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = dYs[(ti + 16 * i) * SP + p];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Xt[p * PADQ + tj + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) r[i][j] += a[i] * b[j];
      }
      // Z[t,q] = W[t,q] G[t,q] = d loss / d (cs_t - cs_q):  d cs_t += sum_q Z,  d cs_q -= sum_t Z
      float zrow[4] = {0.f, 0.f, 0.f, 0.f}, zcol[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t = ti + 16 * i, q = tj + 16 * j;
          const float e = (q <= t) ? __expf(s_cs[t] - s_cs[q]) : 0.f;
          const float g = Gs[t * SQ + q];
          const float wv = r[i][j] * e * s_dt[q];
          Ks[t * SQ + q] = g * e;
          Ws[t * SQ + q] = wv;
          zrow[i] += wv * g; zcol[j] += wv * g;
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = zrow[i];                                           // the 16 lanes of a row group share t
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tj == 0) atomicAdd(&s_dcs[ti + 16 * i], v);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&s_dcs[tj + 16 * j], -zcol[j]);
    }
    __syncthreads();
    const float* Sin = states + (((long long)db * nc + c) * H + h) * (SN * SP);
    const float* Gst = gstates + (((long long)db * nc + c) * H + h) * (SN * SP);
    // ---- du (rows q = ti+16i, cols p = tj+16j)
    {
      float du1[4][4], du2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { du1[i][j] = 0.f; du2[i][j] = 0.f; }
      for (int t = 0; t < SQ; ++t) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Ks[t * SQ + ti + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = dYs[t * SP + tj + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) du1[i][j] += a[i] * b[j];
      }
      for (int n = 0; n < SN; ++n) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Bn[(ti + 16 * i) * SN + n];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = __ldg(Gst + n * SP + tj + 16 * j);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) du2[i][j] += a[i] * b[j];
      }
      float dsum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = ti + 16 * i;
        const float eq = __expf(cs_last - s_cs[q]);
        float dux = 0.f;
        if (q < qv) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int p = tj + 16 * j;
            const float du = du1[i][j] + eq * du2[i][j];
            const float xv = Xt[p * PADQ + q];
            const float dv = dYs[q * SP + p];
            dxc[(row0 + q) * di + h * SP + p] = from_f<T>(s_dt[q] * du + Dh * dv);
            dux += du * xv;
            dsum += dv * xv;
          }
        }
        // reduce over the 16 lanes that share this row
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dux += __shfl_xor_sync(0xffffffffu, dux, o);
        if (tj == 0) s_ddtx[q] = dux;
      }
      dsum = block_sum(dsum, s_red);
      if (tid == 0) atomicAdd(dD + dir * H + h, dsum);
    }
    // ---- dC1 / dB1 (rows wt+8i, cols n = wn+32j)
    for (int k = 0; k < SQ; ++k) {
      float aw[8], awt[8], bb[4], bc[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { aw[i] = Ws[(wt + 8 * i) * SQ + k]; awt[i] = Ws[k * SQ + wt + 8 * i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { bb[j] = Bn[k * SN + wn + 32 * j]; bc[j] = Cn[k * SN + wn + 32 * j]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { accC[i][j] += aw[i] * bb[j]; accB[i][j] += awt[i] * bc[j]; }
    }
    // ---- dC2^T / dB2^T (rows n = ti+16i (i<8), cols t = tj+16j), scaled per head
    {
      float c2[8][4], b2[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c2[i][j] = 0.f; b2[i][j] = 0.f; }
      for (int p = 0; p < SP; ++p) {
        float as[8], ag[8], bd[4], bx[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { as[i] = __ldg(Sin + (ti + 16 * i) * SP + p); ag[i] = __ldg(Gst + (ti + 16 * i) * SP + p); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { bd[j] = dYt[p * PADQ + tj + 16 * j]; bx[j] = Xt[p * PADQ + tj + 16 * j]; }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) { c2[i][j] += as[i] * bd[j]; b2[i][j] += ag[i] * bx[j]; }
      }
      float dlast = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int t = tj + 16 * j;
        const float et = __expf(s_cs[t]);
        const float eq = __expf(cs_last - s_cs[t]) * s_dt[t];
        float yoff = 0.f, sloc = 0.f;                               // <dy_t, Yoff_t> and <dS_c, its q-th term>
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float vc = et * c2[i][j], vb = eq * b2[i][j];
          accC2[i][j] += vc; accB2[i][j] += vb;
          yoff += vc * Cn[t * SN + ti + 16 * i];
          sloc += vb * Bn[t * SN + ti + 16 * i];
        }
        atomicAdd(&s_dcs[t], yoff - sloc);                          // Yoff ~ e^{cs_t};  S_c term ~ e^{cs_last - cs_t}
        dlast += sloc;
      }
      // inter-chunk decay: S_in[c+1] = e^{cs_last} S_in[c] + S_c
      float dot = 0.f;
      for (int i = tid; i < SN * SP; i += ST) dot += __ldg(Gst + i) * __ldg(Sin + i);
      dlast += __expf(cs_last) * dot;
      dlast = block_sum(dlast, s_red);
      if (tid == 0) s_dcs[SQ - 1] += dlast;
    }
    __syncthreads();
    // d(dt_s A) = sum_{t >= s} d cs_t (reverse inclusive cumsum inside the chunk)
    if (tid < 32) {
      float v0 = s_dcs[SQ - 1 - tid], v1 = s_dcs[31 - tid];          // lane 0 holds the latest time of each half
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u0 = __shfl_up_sync(0xffffffffu, v0, o), u1 = __shfl_up_sync(0xffffffffu, v1, o);
        if (tid >= o) { v0 += u0; v1 += u1; }
      }
      v1 += __shfl_sync(0xffffffffu, v0, 31);
      const int s0 = SQ - 1 - tid, s1 = 31 - tid;
      float accA = v0 * s_dt[s0] + v1 * s_dt[s1];
      if (s0 < qv) ddt[(row0 + s0) * H + h] = v0 * A + s_ddtx[s0];
      if (s1 < qv) ddt[(row0 + s1) * H + h] = v1 * A + s_ddtx[s1];
      accA = warp_sum(accA);
      if (tid == 0) atomicAdd(dA_log + dir * H + h, accA * A);
    }
  }
  // ---- write dB | dC for this chunk.  The (t, n)-mapped and the (n, t)-mapped register tiles are summed in
  // shared memory (Bn / Cn are dead by now; every element has one owner per pass) and stored once, coalesced.
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Bn[(wt + 8 * i) * SN + wn + 32 * j] = accB[i][j];
      Cn[(wt + 8 * i) * SN + wn + 32 * j] = accC[i][j];
    }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Bn[(tj + 16 * j) * SN + ti + 16 * i] += accB2[i][j];
      Cn[(tj + 16 * j) * SN + ti + 16 * i] += accC2[i][j];
    }
  __syncthreads();
  for (int i = tid; i < SQ * SN / 4; i += ST) {
    const int t = i / (SN / 4), n4 = (i % (SN / 4)) * 4;
    if (t < qv) {
      stv<T, 4>(dBC + (row0 + t) * (2 * SN) + n4, Bn + t * SN + n4);
      stv<T, 4>(dBC + (row0 + t) * (2 * SN) + SN + n4, Cn + t * SN + n4);
    }
  }
}

}  // namespace hnb

using namespace hnb;

static size_t ssd_states_floats(int ndir, int B, int L, int H) {
  const int nc = cdiv(L, SQ);
  return (size_t)ndir * B * nc * H * SN * SP;
}

extern "C" int hnb_ssd_chunk(void) { return SQ; }

long long hnb_ssd_tc_ws_bytes(int ndir, int B, int L, int H);

extern "C" long long hnb_ssd_ws_bytes(int ndir, int B, int L, int di, int N, int H) {
  (void)di; (void)N;
  const int nc = cdiv(L, SQ);
  const size_t fl = ssd_states_floats(ndir, B, L, H) + (size_t)ndir * B * H * nc + 2 * (size_t)ndir * B * L * H + 64;
  const long long exact = (long long)(fl * sizeof(float)), tc = hnb_ssd_tc_ws_bytes(ndir, B, L, H);
  return exact > tc ? exact : tc;                       // one size serves both implementations
}

static int ssd_check(const char* who, int ndir, int B, int L, int di, int N, int H) {
  if (!(ndir >= 1 && ndir <= 2 && B > 0 && L > 0 && H > 0)) { set_error("%s: bad sizes", who); return HNB_ERR_INVALID_ARG; }
  if (N != SN || di != H * SP) {
    set_error("%s: only d_state=128, headdim=64 are built (got N=%d, di=%d, H=%d)", who, N, di, H);
    return HNB_ERR_UNSUPPORTED;
  }
  return HNB_OK;
}

int hnb_ssd_fwd_tc(const void* xconv, const float* dt, const float* A_log, const float* Dskip, int ndir, int B, int L,
                   int di, int N, int H, void* y, void* states, void* stream, int variant);
int hnb_ssd_bwd_tc(const void* dy, const void* xconv, const void* y, const float* dt, const float* A_log,
                   const float* Dskip, const void* states, int ndir, int B, int L, int di, int N, int H, void* dxc,
                   void* dBC, int dbc_parts, float* ddt, float* dA_log, float* dD, void* ws2, void* stream, int variant);
int hnb_ssd_dbc_parts_tc(int ndir, int B, int L, int H, int variant);

extern "C" int hnb_ssd_dbc_parts(int ndir, int B, int L, int H, int impl) {
  return (impl == 1 || impl == 3 || impl == 4 || impl == 5) ? hnb_ssd_dbc_parts_tc(ndir, B, L, H, impl == 3 ? 1 : 0) : 1;
}

// chunk_state has no per-group recomputation (only the 32 KB B tile is staged again): small groups fill the SMs
static int ssd_heads_per_group_state(int H) {
  static const int env = [] { const char* e = getenv("HNB_SSD_EXACT_HPG_STATE"); return e ? atoi(e) : 0; }();
  const int want = env > 0 ? env : 4;
  return want < H ? want : H;
}

static int ssd_heads_per_group(int H) {
  static const int env = [] { const char* e = getenv("HNB_SSD_EXACT_HPG"); return e ? atoi(e) : 0; }();
  return env > 0 && env < H ? env : H;      // measured (40 rows, B200): 16 / 4 / 1 heads per group = 555 / 581 / 724 us (main stack), 748 / 731 / 972 (outer)
}

template <typename T>
static int ssd_fwd_impl(const T* xconv, const float* dt, const float* A_log, const float* Dskip, int ndir, int B,
                        int L, int di, int H, T* y, float* ws, cudaStream_t st) {
  const int nc = cdiv(L, SQ), C = di + 2 * SN;
  float* states = ws;
  float* decay = ws + ssd_states_floats(ndir, B, L, H);
  const size_t sm1 = (SQ * SN + SQ * SP + 2 * SQ) * sizeof(float);
  const size_t sm3 = SCAN_SMEM_FLOATS * sizeof(float);
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_chunk_state_kernel<T, 0>, (int)sm1));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_chunk_scan_kernel<T>, (int)sm3));
  // Grid (head groups, rows, chunks): chunks slowest, so the short last chunk of every row runs in the tail of the launch
  // (627 -> 555 us on the main stack by itself).  Splitting the heads of a (row, chunk) into groups (HNB_SSD_EXACT_HPG) makes
  // the grid several waves deep but re-stages B | C and recomputes G = C B^T per group: no gain measured, one group by default.
  const int hpg = ssd_heads_per_group(H);
  const int hpg_s = ssd_heads_per_group_state(H);
  dim3 grid(cdiv(H, hpg), ndir * B, nc);
  ssd_chunk_state_kernel<T, 0><<<dim3(cdiv(H, hpg_s), ndir * B, nc), ST, sm1, st>>>(xconv, C, xconv, C, di, dt, A_log, ndir, B, L, H, nc, hpg_s, states, decay);
  HNB_LAUNCH_CHECK("ssd_chunk_state");
  const long long total4 = (long long)ndir * B * H * (SN * SP / 4);
  ssd_state_pass_kernel<false><<<cdiv(total4, 256), 256, 0, st>>>(states, decay, H, nc, total4);
  HNB_LAUNCH_CHECK("ssd_state_pass");
  ssd_chunk_scan_kernel<T><<<grid, ST, sm3, st>>>(xconv, C, di, dt, A_log, Dskip, states, ndir, B, L, H, nc, hpg, y);
  HNB_LAUNCH_CHECK("ssd_chunk_scan");
  return HNB_OK;
}

extern "C" int hnb_ssd_fwd(const void* xconv, int dtype, const float* dt, const float* A_log, const float* Dskip,
                           int ndir, int B, int L, int di, int N, int H, void* y, void* states, int impl,
                           void* stream) {
  HNB_CHECK_ARG(xconv && dt && A_log && Dskip && y && states, "ssd_fwd: null pointer");
  int rc = ssd_check("ssd_fwd", ndir, B, L, di, N, H);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl >= 1 && impl <= 5) {     // 1, 3: by shape; 2 / 4: the persistent one- / two-CTA-per-SM kernels; 5: states pass + scan
    HNB_CHECK_ARG(dtype == HNB_BF16, "ssd_fwd: the tcgen05 path takes bf16 activations");
    return hnb_ssd_fwd_tc(xconv, dt, A_log, Dskip, ndir, B, L, di, N, H, y, states, stream,
                          impl == 2 ? 1 : impl == 4 ? 2 : impl == 5 ? 3 : 0);
  }
  if (dtype == HNB_BF16)
    return ssd_fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)xconv, dt, A_log, Dskip, ndir, B, L, di, H,
                                       (__nv_bfloat16*)y, (float*)states, st);
  if (dtype == HNB_F32)
    return ssd_fwd_impl<float>((const float*)xconv, dt, A_log, Dskip, ndir, B, L, di, H, (float*)y, (float*)states, st);
  set_error("ssd_fwd: unsupported dtype");
  return HNB_ERR_INVALID_ARG;
}

template <typename T>
static int ssd_bwd_impl(const T* dy, const T* xconv, const T* y, const float* dt, const float* A_log,
                        const float* Dskip, const float* ws, int ndir, int B, int L, int di, int H, T* dxc, T* dBC,
                        float* ddt, float* dA_log, float* dD, float* ws2, cudaStream_t st) {
  const int nc = cdiv(L, SQ), C = di + 2 * SN;
  const size_t nst = ssd_states_floats(ndir, B, L, H);
  const float* states = ws;
  const float* decay = ws + nst;
  float* gstates = ws2;
  const size_t sm1 = (SQ * SN + SQ * SP + 2 * SQ) * sizeof(float);
  const size_t smb = (2 * SQ * SN + SN * PADQ + SQ * SQ + SQ * SP + 2 * SQ * SQ + 4 * SQ + 32) * sizeof(float);
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_chunk_state_kernel<T, 1>, (int)sm1));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_chunk_kernel<T>, (int)smb));
  dim3 grid(nc, ndir * B);
  const int hpg = ssd_heads_per_group_state(H);
  ssd_chunk_state_kernel<T, 1><<<dim3(cdiv(H, hpg), ndir * B, nc), ST, sm1, st>>>(dy, di, xconv, C, di, dt, A_log, ndir, B, L, H, nc, hpg, gstates, nullptr);
  HNB_LAUNCH_CHECK("ssd_bwd_dstate");
  const long long total4 = (long long)ndir * B * H * (SN * SP / 4);
  ssd_state_pass_kernel<true><<<cdiv(total4, 256), 256, 0, st>>>(gstates, decay, H, nc, total4);
  HNB_LAUNCH_CHECK("ssd_state_pass_rev");
  (void)y;
  ssd_bwd_chunk_kernel<T><<<grid, ST, smb, st>>>(dy, xconv, C, di, dt, A_log, Dskip, states, gstates, ndir, B, L, H, nc,
                                                 dxc, dBC, ddt, dA_log, dD);
  HNB_LAUNCH_CHECK("ssd_bwd_chunk");
  return HNB_OK;
}

extern "C" int hnb_ssd_bwd(const void* dy, const void* xconv, const void* y, int dtype, const float* dt,
                           const float* A_log, const float* Dskip, const void* states, int ndir, int B, int L, int di,
                           int N, int H, void* dxc, void* dBC, int dbc_parts, float* ddt, float* dA_log, float* dD,
                           void* ws2, int impl, void* stream) {
  HNB_CHECK_ARG(dy && xconv && y && dt && A_log && Dskip && states && dxc && dBC && ddt && dA_log && dD && ws2,
                "ssd_bwd: null pointer");
  int rc = ssd_check("ssd_bwd", ndir, B, L, di, N, H);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1 || impl == 3 || impl == 4 || impl == 5) {   // 1, 4, 5: fused dx + dB/dC kernel; 3: the three-kernel backward (kept under test)
    HNB_CHECK_ARG(dtype == HNB_BF16, "ssd_bwd: the tcgen05 path takes bf16 activations");
    return hnb_ssd_bwd_tc(dy, xconv, y, dt, A_log, Dskip, states, ndir, B, L, di, N, H, dxc, dBC, dbc_parts, ddt, dA_log,
                          dD, ws2, stream, impl == 3 ? 1 : 0);
  }
  HNB_CHECK_ARG(dbc_parts == 1, "ssd_bwd: the CUDA-core path writes one dBC part");
  if (dtype == HNB_BF16)
    return ssd_bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)xconv, (const __nv_bfloat16*)y,
                                       dt, A_log, Dskip, (const float*)states, ndir, B, L, di, H, (__nv_bfloat16*)dxc,
                                       (__nv_bfloat16*)dBC, ddt, dA_log, dD, (float*)ws2, st);
  if (dtype == HNB_F32)
    return ssd_bwd_impl<float>((const float*)dy, (const float*)xconv, (const float*)y, dt, A_log, Dskip,
                               (const float*)states, ndir, B, L, di, H, (float*)dxc, (float*)dBC, ddt, dA_log, dD, (float*)ws2, st);
  set_error("ssd_bwd: unsupported dtype");
  return HNB_ERR_INVALID_ARG;
}
