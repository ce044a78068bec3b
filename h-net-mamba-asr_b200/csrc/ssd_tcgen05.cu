// SSD (Mamba-2 selective scan) on tcgen05 tensor cores (impl 1): bf16 operands staged by TMA, fp32
// accumulators in TMEM, chunk length 128.  Scan order in, scan order out (see ssd_kernels.cu for the maths).
//
// Forward: ONE persistent kernel.  A CTA owns a (row, head) pair and walks its chunks in order, carrying the
// 128x64 state in registers (fp32) and in shared memory (bf16, the next chunk's MMA operand):
//     G    = C B^T                        128x128x128   (TMEM cols   0..127)
//     M    = bf16( G o L o dt )           epilogue 1    -> smem, K-major A operand
//     Yd   = M X                          128x64x128    (TMEM cols 128..191)   X  : MN-major B operand (as TMA wrote it)
//     Yo   = C S_in                       128x64x128    (TMEM cols 192..255)   S  : MN-major B operand
//     dS   = B^T (w o X)                  128x64x128    (TMEM cols 256..319)   B^T: MN-major A operand
//     y    = Yd + e^{cs} Yo + D x         epilogue 2    -> global (bf16)
//     S    = e^{cs_last} S + dS           epilogue 3    -> registers, smem (bf16), global (bf16, for the backward)
//
// Backward (three kernels):
//   1. dstate : per (row, head), chunks in REVERSE order:  Gst[c] = d loss / d (state leaving chunk c)
//   2. dx     : per (row, chunk, head): du = K^T dY + e_q B Gst  ->  dx, ddt, dA_log, dD
//   3. dbc    : per (row, chunk), heads accumulated in TMEM:   dC += W B + (e^{cs} dY) S_in^T,
//                                                              dB += W^T C + (w X) Gst^T
// with  K = (C B^T) o L,  W = (dY X^T) o L o dt,  L[t,q] = e^{cs_t - cs_q} (q <= t).
//
// These kernels are paced by the CUDA-core epilogues between the MMAs (decay matrix, casts, row dots), not by the
// tensor pipe, so: 8-16 warps per CTA (warp w reads TMEM lane quarter w%4, column group w/4; the last warp issues TMA
// and MMAs while the others build tables), every table in the .shared address space, the decay matrix L factored per 32-row block through a reference
// point (both factors <= 1: no overflow; only the diagonal 32x32 blocks pay one exp per element), the per-chunk
// tables of the NEXT step built while the current step's MMAs run, and TMA for the next step issued as soon as the
// current operands are consumed.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"
#include "ssd_tc.cuh"

namespace hnb {
namespace {

// phase timers exist only in the DBG instantiations (HNB_SSD_DEBUG=1): they cost ten live registers otherwise
#define HNB_CLK() (DBG ? clock64() : 0LL)


// The tables of every (row, head, chunk), built ONCE per forward and kept in the workspace: the forward and the three
// backward kernels fetch them with one 1-D bulk copy next to their TMA tile loads, so no scan / exp / barrier for
// them sits on a kernel's critical path any more.
__global__ void __launch_bounds__(TQ)
ssd_tables_kernel(const float* __restrict__ dt, const float* __restrict__ A_log, float* __restrict__ tables, int B, int L,
                  int H, int nc, int* __restrict__ work_counter) {
  pdl_enter();
  __shared__ float tab[TAB_FLOATS];
  if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0;          // the forward kernel's dynamic work queue
  const int item = blockIdx.x;                                        // ((db * H) + h) * nc + c
  const int c = item % nc, h = (item / nc) % H, db = item / (nc * H), dir = db / B;
  const int q0 = c * TQ;
  build_tables(dt + ((long long)db * L + q0) * H + h, H, min(TQ, L - q0), -__expf(A_log[dir * H + h]), tab, TQ);
  float* out = tables + (long long)item * TAB_FLOATS;
  for (int i = threadIdx.x; i < 9 * TQ; i += TQ) out[i] = tab[i];
}


// ===================================================================================================
// forward
// ===================================================================================================
constexpr int OFF_C = 0;                // 2 blocks: n in [0,64) | [64,128)
constexpr int OFF_B = OFF_C + 2 * HALF;
constexpr int OFF_X = OFF_B + 2 * HALF; // 2 buffers (chunk parity)
constexpr int OFF_XW = OFF_X + 2 * HALF;
constexpr int OFF_M = OFF_XW + HALF;    // 2 blocks: s in [0,64) | [64,128)
constexpr int OFF_S = OFF_M + 2 * HALF;
constexpr int OFF_TAB = OFF_S + HALF;   // two table buffers (current / next chunk)
constexpr int OFF_BAR = OFF_TAB + 2 * TAB_BYTES;
constexpr int FWD_SMEM = OFF_BAR + 64 + 1024;


template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
ssd_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const FwdParams p) {
  constexpr int NCG = NT / 128;           // column groups: warp w -> TMEM lane quarter w%4, column group w/4
  static_assert(NCG == 4, "the balanced score-tile split assumes 16 warps");
  constexpr int NB = 4 / NCG;             // 32-column blocks (of 128) and 16-column blocks (of 64) per thread
  constexpr int NTAB = NT - 32;           // lane 0 of the last warp issues TMA and MMAs
  extern __shared__ __align__(1024) uint8_t smem_raw[];   // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base = smem_raw;                               // (no integer round trip: keeps the .shared address space)
  uint8_t* sC = base + OFF_C; uint8_t* sB = base + OFF_B; uint8_t* sX = base + OFF_X; uint8_t* sXw = base + OFF_XW;
  uint8_t* sM = base + OFF_M; uint8_t* sS = base + OFF_S;
  float* tabs = reinterpret_cast<float*>(base + OFF_TAB);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(base + OFF_BAR);
  uint64_t* bar_g = bar_load + 1;
  uint64_t* bar_y = bar_load + 2;
  uint64_t* bar_free = bar_load + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2;
  const int row = lq * 32 + lane;
  const bool issuer = tid == NTAB;                   // lane 0 of the last warp: MMA issue
  // TMA issue: lane 0 of the warp that owns the top-right 32x32 block of the score matrix (all zeros: the warp with
  // the least epilogue work), so waiting for the operands to be released costs nobody else time
  const bool loader = tid == 128 * (NCG - 1);

  if (tid == 0) {
    umma::prefetch_tmap(&tmX);
    umma::mbar_init(bar_load, 1); umma::mbar_init(bar_g, 1); umma::mbar_init(bar_y, 1); umma::mbar_init(bar_free, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);

  const int H = p.H, L = p.L, di = p.di, nc = p.nc;
  const int n_items = p.ndirB * H;
  constexpr uint32_t idesc_g = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_y = umma::make_idesc_bf16(128, 64, 0, 1);
  constexpr uint32_t idesc_s = umma::make_idesc_bf16(128, 64, 1, 1);

  auto issue_load = [&](int item, int c, int buf) {                 // loader only
    int db, h; p.dH.divmod(item, db, h);
    umma::mbar_expect_tx(bar_load, 5 * HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)(item * nc + c) * TAB_FLOATS, TAB_BYTES, bar_load);
    umma::tma_load_3d(sC, &tmX, bar_load, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_load, di + TN + 64, c * TQ, db);
    umma::tma_load_3d(sB, &tmX, bar_load, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_load, di + 64, c * TQ, db);
    umma::tma_load_3d(sX + buf * HALF, &tmX, bar_load, h * TP, c * TQ, db);
  };

  uint32_t seq = 0;                                                   // (item, chunk) sequence number of this CTA
  if (blockIdx.x < n_items && loader) issue_load(blockIdx.x, 0, 0);

  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    int db, h; p.dH.divmod(it, db, h);
    const int dir = p.dB.div(db);
    const float Dh = p.Dskip[dir * H + h];
    float Sreg[16 * NB];
#pragma unroll
    for (int j = 0; j < 16 * NB; ++j) Sreg[j] = 0.f;
    for (int i = tid; i < HALF / 16; i += NT) reinterpret_cast<uint4*>(sS)[i] = make_uint4(0, 0, 0, 0);

    for (int c = 0; c < nc; ++c, ++seq) {
      const int buf = seq & 1;
      const uint32_t par = seq & 1;
      const float* tab = tabs + buf * TAB_FLOATS;
      const float* s_dt = tab + TQ; const float* s_w = tab + 2 * TQ; const float* s_ecs = tab + 3 * TQ;
      const int q0 = c * TQ, qv = min(TQ, L - q0);
      const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;           // 32-row blocks / 16-row k-steps holding valid frames
      const bool last_chunk = c == nc - 1;                             // nothing consumes the state leaving it
      const long long row0 = (long long)db * L + q0;
      int nit = it, ncn = c + 1;                                       // the step after this one
      if (ncn == nc) { nit = it + gridDim.x; ncn = 0; }
      {   // state entering this chunk, kept for the backward (thread = state row n, 64/NCG of the 64 columns)
        __nv_bfloat16* sg = p.states + ((((long long)db * H + h) * nc + c) * TN + row) * TP + 16 * NB * cg;
#pragma unroll
        for (int k = 0; k < 2 * NB; ++k) *reinterpret_cast<uint4*>(sg + 8 * k) = pack8(Sreg + 8 * k);
      }
      const long long tk0 = HNB_CLK();
      umma::mbar_wait(bar_load, par);
      const long long tk1 = HNB_CLK();
      if (issuer) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          umma::mma_bf16_ss(tmem + 0, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024), idesc_g, kb > 0);
        }
        umma::mma_commit(bar_g);
      }
      // Xw = w_s * X (same swizzled positions: the swizzle permutes 16-byte chunks inside a row only)
      if (!last_chunk) for (int i = tid; i < TQ * 8; i += NT) {
        const float w = s_w[i >> 3];
        float v[8];
        unpack8(reinterpret_cast<const uint4*>(sX + buf * HALF)[i], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= w;
        reinterpret_cast<uint4*>(sXw)[i] = pack8(v);
      }
      umma::fence_async_smem();
      umma::tc_fence_before();
      __syncthreads();                                                 // Xw (and the state in smem) visible to the tensor core
      if (issuer) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // Yo = C S_in
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          umma::mma_bf16_ss(tmem + 192, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sS) + kb * 2048, 1024, 1024), idesc_y, kb > 0);
        }
        if (!last_chunk) {
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {                             // dS = B^T (w o X)   (k = time: valid frames only)
            if (kb < nkb)
              umma::mma_bf16_ss(tmem + 256, umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024),
                                umma::make_smem_desc(umma::smem_u32(sXw) + kb * 2048, 1024, 1024), idesc_s, kb > 0);
          }
        }
        umma::mma_commit(bar_free);                                    // C, B, Xw, S_in released when these retire
      }
      if (loader && nit < n_items) {                                   // next step's tiles land during this step's epilogues
        umma::mbar_wait(bar_free, par);
        const long long tl0 = HNB_CLK();
        issue_load(nit, ncn, buf ^ 1);
        if (DBG && p.dbg && blockIdx.x == 0) {                                // debug only: how long the TMA group takes to land
          umma::mbar_wait(bar_load, par ^ 1);
          p.dbg[5] += HNB_CLK() - tl0; p.dbg[7] += tl0 - tk1;
        }
      }
      umma::mbar_wait(bar_g, par);
      umma::tc_fence_after();
      const long long tk2 = HNB_CLK();
      // ---- epilogue 1: M[t,s] = G[t,s] e^{cs_t - cs_s} dt_s (s <= t), bf16, K-major swizzled
      if ((row >> 5) < nblk) {                                         // padding rows: their M rows only feed unused y rows
        const int t = row, I = t >> 5;
        const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
        float g[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k <= I) umma::tmem_ld8(t_lane + (uint32_t)(32 * k + 8 * cg), g[k]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s0 = 32 * k + 8 * cg;
          if (k <= I) {
            float l[8], d8[8];
            decay8(l, t, I, k, s0, cs_t, e_ref, tab);
            load8(d8, s_dt + s0);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[k][j] *= l[j] * d8[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[k][j] = 0.f;
          }
          *reinterpret_cast<uint4*>(sM + (k >> 1) * HALF + swz(t, 4 * (k & 1) + cg)) = pack8(g[k]);
        }
      }
      umma::fence_async_smem();
      umma::tc_fence_before();
      __syncthreads();
      const long long tk3 = HNB_CLK();
      if (issuer) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // Yd = M X   (k = time: valid frames only)
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          if (kb < nkb)
            umma::mma_bf16_ss(tmem + 128, umma::make_smem_desc(umma::smem_u32(sM) + o, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sX + buf * HALF) + kb * 2048, 1024, 1024), idesc_y, kb > 0);
        }
        umma::mma_commit(bar_y);
      }
      umma::mbar_wait(bar_y, par);
      umma::tc_fence_after();
      const long long tk4 = HNB_CLK();
      // ---- epilogue 2: y = Yd + e^{cs_t} Yo + D x          (thread = row t, 64/NCG of the 64 columns)
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const int t = row, c16 = NB * cg + bb;                         // 16-column block of the 64
        if ((t >> 5) >= nblk) break;
        float yd[16], yo[16];
        umma::tmem_ld16(t_lane + 128u + 16u * c16, yd);
        umma::tmem_ld16(t_lane + 192u + 16u * c16, yo);
        umma::tmem_ld_wait();
        if (t < qv) {
          const float ecs = s_ecs[t];
          __nv_bfloat16* yg = p.y + (row0 + t) * di + h * TP + 16 * c16;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            float x[8], o[8];
            unpack8(*reinterpret_cast<const uint4*>(sX + buf * HALF + swz(t, 2 * c16 + k)), x);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = yd[8 * k + e] + ecs * yo[8 * k + e] + Dh * x[e];
            *reinterpret_cast<uint4*>(yg + 8 * k) = pack8(o);
          }
        }
      }
      // ---- epilogue 3: S = e^{cs_last} S + dS  (thread = state row n)
      if (!last_chunk) {
        const float decay = s_ecs[TQ - 1];
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          const int c16 = NB * cg + bb;
          float ds[16];
          umma::tmem_ld16(t_lane + 256u + 16u * c16, ds);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) Sreg[16 * bb + j] = decay * Sreg[16 * bb + j] + ds[j];
          *reinterpret_cast<uint4*>(sS + swz(row, 2 * c16)) = pack8(Sreg + 16 * bb);
          *reinterpret_cast<uint4*>(sS + swz(row, 2 * c16 + 1)) = pack8(Sreg + 16 * bb + 8);
        }
      }
      umma::fence_async_smem();
      umma::tc_fence_before();
      __syncthreads();
      if (DBG && p.dbg && blockIdx.x == 0 && tid == 0) {
        const long long tk5 = HNB_CLK();
        p.dbg[0] += tk1 - tk0; p.dbg[1] += tk2 - tk1; p.dbg[2] += tk3 - tk2; p.dbg[3] += tk4 - tk3;
        p.dbg[4] += tk5 - tk4; p.dbg[6] += 1;
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ===================================================================================================
// forward, two CTAs per SM
// ===================================================================================================
// The one-CTA forward above is latency-bound: every chunk step is a chain  TMA -> G -> epilogue -> Yd -> epilogue ->
// barrier  with the tensor pipe ~15 % and the issue slots ~30 % busy.  Two independent chains per SM hide each other's
// waits, which needs <= 113 KB of shared memory and <= 256 TMEM columns per CTA:
//   * the bf16 score tile M goes to TMEM (packed pairs) and is the A operand of Yd = M X straight from there
//     (tcgen05.mma with A in TMEM): no 32 KB smem tile, no proxy fence on that hand-off;
//   * w o X overwrites X in place once Yd has retired (x for the D x term is re-read from global: L2 hits);
//   * one X buffer, loads for the next step issued as their buffers retire (C | tables after Yo, B | X after dS);
//   * TMEM: G [0,128) -> reused by Yo [0,64) and dS [64,128) once epilogue 1 has read G;  M [128,192);  Yd [192,256).
// 256 threads: warp w reads TMEM lane quarter w % 4 and column half w / 4.
constexpr int F2_THREADS = 256;
constexpr int F2_OFF_C = 0, F2_OFF_B = 2 * HALF, F2_OFF_X = 4 * HALF, F2_OFF_S = 5 * HALF, F2_OFF_TAB = 6 * HALF;
constexpr int F2_OFF_BAR = F2_OFF_TAB + 2 * TAB_BYTES;
constexpr int F2_SMEM = F2_OFF_BAR + 64 + 1024;
static_assert(2 * (F2_SMEM + 1024) <= 228 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(F2_THREADS, 2)
ssd_fwd2_tc_kernel(const __grid_constant__ CUtensorMap tmX, const FwdParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + F2_OFF_C; uint8_t* sB = base + F2_OFF_B; uint8_t* sX = base + F2_OFF_X; uint8_t* sS = base + F2_OFF_S;
  float* tabs = reinterpret_cast<float*>(base + F2_OFF_TAB);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(base + F2_OFF_BAR);
  uint64_t* bar_g = bar_load + 1;
  uint64_t* bar_y = bar_load + 2;
  uint64_t* bar_free = bar_load + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 4);
  volatile int* s_next = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [2] next item of this CTA, by item parity
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2;                 // TMEM lane quarter, column half
  const int row = lq * 32 + lane;
  const bool issuer = tid == 0;                            // warps 0 and 4 own the lightest score-tile rows
  const bool loader = tid == 128;
  if (tid == 0) {
    umma::prefetch_tmap(&tmX);
    umma::mbar_init(bar_load, 2);                          // two load groups per step, each one arrive.expect_tx
    umma::mbar_init(bar_g, 1); umma::mbar_init(bar_y, 1); umma::mbar_init(bar_free, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  constexpr uint32_t TM_G = 0, TM_YO = 0, TM_DS = 64, TM_M = 128, TM_YD = 192;
  const int H = p.H, L = p.L, di = p.di, nc = p.nc, C = di + 2 * TN;
  const int n_items = p.ndirB * H;
  constexpr uint32_t idesc_g = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_y = umma::make_idesc_bf16(128, 64, 0, 1);
  constexpr uint32_t idesc_s = umma::make_idesc_bf16(128, 64, 1, 1);

  auto load_a = [&](int item, int c, int buf) {            // C | tables: free once G and Yo have retired
    int db, h; p.dH.divmod(item, db, h);
    umma::mbar_expect_tx(bar_load, 2 * HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)(item * nc + c) * TAB_FLOATS, TAB_BYTES, bar_load);
    umma::tma_load_3d(sC, &tmX, bar_load, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_load, di + TN + 64, c * TQ, db);
  };
  auto load_b = [&](int item, int c) {                     // B | X: free once dS has retired
    int db, h; p.dH.divmod(item, db, h);
    umma::mbar_expect_tx(bar_load, 3 * HALF);
    umma::tma_load_3d(sB, &tmX, bar_load, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_load, di + 64, c * TQ, db);
    umma::tma_load_3d(sX, &tmX, bar_load, h * TP, c * TQ, db);
  };

  uint32_t seq = 0, fseq = 0, iseq = 0;                    // chunk steps / dS steps / items of this CTA
  if (blockIdx.x < n_items && loader) { load_a(blockIdx.x, 0, 0); load_b(blockIdx.x, 0); }

  // Items are handed out round-robin, or (HNB_SSD_FWD_DYNAMIC=1) claimed from a global counter.  Measured at the
  // headline shapes the two are within 2 % of each other (equal-cost items: 960 on 296 CTAs is 4 rounds either way);
  // the counter is for ragged batches, where rows differ in their number of chunks.
  int my_next = n_items;                                   // loader only: the item after this one
  for (int it = blockIdx.x; it < n_items; ++iseq) {
    if (loader)                                            // consumed after the first G wait: the atomic's latency is hidden
      my_next = p.counter ? (int)gridDim.x + atomicAdd(p.counter, 1) : it + (int)gridDim.x;
    int db, h; p.dH.divmod(it, db, h);
    const int dir = p.dB.div(db);
    const float Dh = p.Dskip[dir * H + h];
    float Sreg[32];                                        // thread = state row n, 32 of the 64 columns
#pragma unroll
    for (int j = 0; j < 32; ++j) Sreg[j] = 0.f;
    for (int i = tid; i < HALF / 16; i += F2_THREADS) reinterpret_cast<uint4*>(sS)[i] = make_uint4(0, 0, 0, 0);

    for (int c = 0; c < nc; ++c, ++seq) {
      const int buf = seq & 1;
      const uint32_t par = seq & 1;
      const float* tab = tabs + buf * TAB_FLOATS;
      const float* s_dt = tab + TQ; const float* s_w = tab + 2 * TQ; const float* s_ecs = tab + 3 * TQ;
      const int q0 = c * TQ, qv = min(TQ, L - q0);
      const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;
      const bool last_chunk = c == nc - 1;
      const long long row0 = (long long)db * L + q0;
      int nit = it, ncn = c + 1;                                       // the step after this one (loader only)
      if (ncn == nc) { nit = my_next; ncn = 0; }
      {   // state entering this chunk, kept for the backward
        __nv_bfloat16* sg = p.states + ((((long long)db * H + h) * nc + c) * TN + row) * TP + 32 * cg;
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(sg + 8 * k) = pack8(Sreg + 8 * k);
      }
      umma::mbar_wait(bar_load, par);
      if (issuer) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // G = C B^T
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          umma::mma_bf16_ss(tmem + TM_G, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024), idesc_g, kb > 0);
        }
        umma::mma_commit(bar_g);
      }
      umma::mbar_wait(bar_g, par);
      umma::tc_fence_after();
      if (loader && c == 0) s_next[iseq & 1] = my_next;
      // ---- epilogue 1: M[t,s] = G[t,s] e^{cs_t - cs_s} dt_s (s <= t) -> bf16 pairs in TMEM (A operand of Yd)
      if ((row >> 5) < nblk) {                                         // padding rows only feed unused y rows
        const int t = row, I = t >> 5;
        const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {                               // two 32-column blocks per TMEM round trip
          float g[2][16];
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            if (2 * kp + kk <= I) umma::tmem_ld16(t_lane + TM_G + (uint32_t)(32 * (2 * kp + kk) + 16 * cg), g[kk]);
          umma::tmem_ld_wait();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int k = 2 * kp + kk, s0 = 32 * k + 16 * cg;
            uint32_t pk[8];
            if (k <= I) {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                float l[8], d8[8];
                decay8(l, t, I, k, s0 + 8 * hf, cs_t, e_ref, tab);
                load8(d8, s_dt + s0 + 8 * hf);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[kk][8 * hf + j] *= l[j] * d8[j];
                const uint4 q = pack8(g[kk] + 8 * hf);
                pk[4 * hf] = q.x; pk[4 * hf + 1] = q.y; pk[4 * hf + 2] = q.z; pk[4 * hf + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = 0u;
            }
            umma::tmem_st8(t_lane + TM_M + (uint32_t)(16 * k + 8 * cg), pk);
          }
        }
        umma::tmem_st_wait();
      }
      umma::fence_async_smem();                                        // the state tile written through the generic proxy
      umma::tc_fence_before();
      __syncthreads();
      if (issuer) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)                                 // Yd = M X   (k = time: valid frames only)
          if (kb < nkb)
            umma::mma_bf16_ts(tmem + TM_YD, tmem + TM_M + 8 * kb,
                              umma::make_smem_desc(umma::smem_u32(sX) + kb * 2048, 1024, 1024), idesc_y, kb > 0);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // Yo = C S_in  (into the columns G has left)
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          umma::mma_bf16_ss(tmem + TM_YO, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sS) + kb * 2048, 1024, 1024), idesc_y, kb > 0);
        }
        umma::mma_commit(bar_y);
      }
      // x of this thread's row for the D x term: the lines TMA fetched a moment ago, still in L2
      uint4 xg[4];
      {
        const uint4* xp = reinterpret_cast<const uint4*>(p.xconv + (row0 + row) * C + h * TP + 32 * cg);
#pragma unroll
        for (int k = 0; k < 4; ++k) xg[k] = row < qv ? __ldg(xp + k) : make_uint4(0, 0, 0, 0);
      }
      umma::mbar_wait(bar_y, par);
      umma::tc_fence_after();
      if (loader && nit < n_items) {
        load_a(nit, ncn, buf ^ 1);
        if (last_chunk) load_b(nit, ncn);                              // no dS step: B and X are free as well
      }
      // ---- epilogue 2: y = Yd + e^{cs_t} Yo + D x
      if ((row >> 5) < nblk) {
        const int t = row;
        const float ecs = s_ecs[t];
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int c16 = 2 * cg + bb;
          float yd[16], yo[16];
          umma::tmem_ld16(t_lane + TM_YD + 16u * c16, yd);
          umma::tmem_ld16(t_lane + TM_YO + 16u * c16, yo);
          umma::tmem_ld_wait();
          if (t < qv) {
            __nv_bfloat16* yg = p.y + (row0 + t) * di + h * TP + 16 * c16;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              float x[8], o[8];
              unpack8(xg[2 * bb + k], x);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = yd[8 * k + e] + ecs * yo[8 * k + e] + Dh * x[e];
              *reinterpret_cast<uint4*>(yg + 8 * k) = pack8(o);
            }
          }
        }
      }
      if (!last_chunk) {
        // Xw = w_s X, in place (Yd has retired; the swizzle permutes 16-byte chunks inside a row only)
        for (int i = tid; i < TQ * 8; i += F2_THREADS) {
          const float w = s_w[i >> 3];
          float v[8];
          unpack8(reinterpret_cast<const uint4*>(sX)[i], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= w;
          reinterpret_cast<uint4*>(sX)[i] = pack8(v);
        }
        umma::fence_async_smem();
        umma::tc_fence_before();
        __syncthreads();                                               // Xw visible; Yo read by every thread
        if (issuer) {
          umma::tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)                               // dS = B^T (w o X)   (k = time: valid frames only)
            if (kb < nkb)
              umma::mma_bf16_ss(tmem + TM_DS, umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024),
                                umma::make_smem_desc(umma::smem_u32(sX) + kb * 2048, 1024, 1024), idesc_s, kb > 0);
          umma::mma_commit(bar_free);
        }
        umma::mbar_wait(bar_free, fseq & 1);
        ++fseq;
        umma::tc_fence_after();
        if (loader && nit < n_items) load_b(nit, ncn);
        // ---- epilogue 3: S = e^{cs_last} S + dS  (thread = state row n)
        const float decay = s_ecs[TQ - 1];
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int c16 = 2 * cg + bb;
          float ds[16];
          umma::tmem_ld16(t_lane + TM_DS + 16u * c16, ds);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) Sreg[16 * bb + j] = decay * Sreg[16 * bb + j] + ds[j];
          *reinterpret_cast<uint4*>(sS + swz(row, 2 * c16)) = pack8(Sreg + 16 * bb);
          *reinterpret_cast<uint4*>(sS + swz(row, 2 * c16 + 1)) = pack8(Sreg + 16 * bb + 8);
        }
      }
      umma::fence_async_smem();
      umma::tc_fence_before();
      __syncthreads();
    }
    it = s_next[iseq & 1];                                 // written at the start of this item, several barriers ago
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// ===================================================================================================
// backward
// ===================================================================================================
struct BwdParams {
  const float* dt; const float* A_log; const float* Dskip;
  const float* tables;           // [ndir*B, H, nc, TAB_FLOATS]  (built by the forward)
  const __nv_bfloat16* xconv;    // [ndir*B*L, di + 2N]  (fused kernel: x and dY rows re-read for epilogue B, L2 hits)
  const __nv_bfloat16* dy;       // [ndir*B*L, di]
  __nv_bfloat16* gstates;        // [ndir*B, H, nc, 128, 64]
  __nv_bfloat16* dxc;            // [ndir*B*L, di]
  __nv_bfloat16* dBC;            // [HG][ndir*B*L, 2N]: one partial sum per head group (summed by the conv backward)
  long long dbc_part_stride;     // elements between the parts
  int HG;                        // head groups of the dB/dC kernel (H % HG == 0)
  float* ddt;                    // [ndir*B*L, H]
  float* dA_log; float* dD;      // [ndir, H]
  int ndirB, B, L, H, di, nc;
  FastDiv dH, dB, dnc;
  // Work-item order of the dx and dB/dC kernels: every full 128-frame chunk first, the partly filled last chunk of
  // each row (cheap: its MMAs and epilogues are trimmed to the valid frames) at the end, so that the static
  // round-robin over CTAs hands out the expensive items evenly and the cheap ones fill the last wave.
  int nfc, n_full;               // full chunks per row, ndirB * nfc
  FastDiv dnfc, dHG;
  int dxHh;                      // dx kernel: heads per work item (C, B and G = C B^T are shared by an item's heads)
  FastDiv dHGx;                  // H / dxHh
  __device__ __forceinline__ void chunk_of(int base, int& db, int& c) const {
    if (base < n_full) dnfc.divmod(base, db, c); else { db = base - n_full; c = nc - 1; }
  }
  // Fused kernel, span schedule: the head-steps of all items, in the order above, form ONE sequence g = base * H + h; CTA k
  // works the contiguous piece [cuts.g[k], cuts.g[k + 1]) of it.  The host places the cuts at equal cost and moves a cut
  // to the item's end where an item would be cut twice, so every item is cut at most once.  dB | dC of a cut item (a sum
  // over its heads) is finished by the CTA holding the item's FIRST heads: the CTA holding the last heads meets the item
  // as its first piece, leaves its fp32 partial sum in fix[k] and raises fix_flags[k]; its left neighbour reaches the
  // item as its LAST piece, long after, and adds the partial sum before it rounds and stores (stream-K style fix-up; all
  // CTAs are resident: one per SM).  fix_flags are cleared by the state-gradient kernel that runs in front.
  float* fix;                    // [G][2][8][4][128] float4
  int* fix_flags;                // [SPAN_MAX_CTAS + 1]
  long long* dbg;                // optional [8] phase-cycle accumulators of CTA 0 (HNB_SSD_DEBUG=1)
};

constexpr int SPAN_MAX_CTAS = 160;
struct SpanCuts { int g[SPAN_MAX_CTAS + 1]; };

// ---- 1. dstate (256 threads, 2 CTAs per SM) -----------------------------------------------------------
constexpr int D1_THREADS = 256;
constexpr int D1_OFF_C = 0, D1_OFF_DY = 2 * HALF, D1_OFF_DYS = 3 * HALF, D1_OFF_TAB = 4 * HALF;
constexpr int D1_OFF_BAR = D1_OFF_TAB + 2 * TAB_BYTES;
constexpr int D1_SMEM = D1_OFF_BAR + 64 + 1024;

__global__ void __launch_bounds__(D1_THREADS, 2)
ssd_bwd_dstate_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                         const BwdParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + D1_OFF_C; uint8_t* sdY = base + D1_OFF_DY; uint8_t* sdYs = base + D1_OFF_DYS;
  float* tabs = reinterpret_cast<float*>(base + D1_OFF_TAB);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(base + D1_OFF_BAR);
  uint64_t* bar_m = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, ch = warp >> 2, row = lq * 32 + lane;
  if (tid == 0) {
    umma::prefetch_tmap(&tmX); umma::prefetch_tmap(&tmDY);
    umma::mbar_init(bar_load, 1); umma::mbar_init(bar_m, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 64);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  const int H = p.H, L = p.L, di = p.di, nc = p.nc, n_items = p.ndirB * H;
  if (blockIdx.x == 0 && tid <= SPAN_MAX_CTAS && p.fix_flags) p.fix_flags[tid] = 0;   // the fused kernel's fix-up counters
  constexpr uint32_t idesc = umma::make_idesc_bf16(128, 64, 1, 1);
  auto issue_load = [&](int item, int c, int buf) {
    int db, h; p.dH.divmod(item, db, h);
    umma::mbar_expect_tx(bar_load, 3 * HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)(item * nc + c) * TAB_FLOATS, TAB_BYTES, bar_load);
    umma::tma_load_3d(sC, &tmX, bar_load, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_load, di + TN + 64, c * TQ, db);
    umma::tma_load_3d(sdY, &tmDY, bar_load, h * TP, c * TQ, db);
  };
  uint32_t seq = 0;
  // chunk 0 needs no step of its own: the gradient w.r.t. the (zero) state entering it is never used
  if (blockIdx.x < n_items && tid == 0 && nc > 1) issue_load(blockIdx.x, nc - 1, 0);
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    int db, h; p.dH.divmod(it, db, h);
    float Grun[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) Grun[j] = 0.f;
    for (int c = nc - 1; c >= 1; --c, ++seq) {
      const uint32_t par = seq & 1;
      const float* s_ecs = tabs + (seq & 1) * TAB_FLOATS + 3 * TQ;
      const int q0 = c * TQ, qv = min(TQ, L - q0);
      const int nkb = (qv + 15) >> 4;
      {   // gradient w.r.t. the state LEAVING this chunk
        __nv_bfloat16* gg = p.gstates + ((((long long)db * H + h) * nc + c) * TN + row) * TP + 32 * ch;
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(gg + 8 * k) = pack8(Grun + 8 * k);
      }
      umma::mbar_wait(bar_load, par);
      for (int i = tid; i < TQ * 8; i += D1_THREADS) {
        const float e = s_ecs[i >> 3];
        float v[8];
        unpack8(reinterpret_cast<const uint4*>(sdY)[i], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= e;
        reinterpret_cast<uint4*>(sdYs)[i] = pack8(v);
      }
      umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads();
      if (tid == 0) {
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)                                 // k = time: valid frames only
          if (kb < nkb)
            umma::mma_bf16_ss(tmem, umma::make_smem_desc(umma::smem_u32(sC) + kb * 2048, HALF, 1024),
                              umma::make_smem_desc(umma::smem_u32(sdYs) + kb * 2048, 1024, 1024), idesc, kb > 0);
        umma::mma_commit(bar_m);
      }
      umma::mbar_wait(bar_m, par);
      umma::tc_fence_after();
      if (tid == 0) {
        int nit = it, ncn = c - 1;
        if (ncn < 1) { nit = it + gridDim.x; ncn = nc - 1; }
        if (nit < n_items) issue_load(nit, ncn, (seq & 1) ^ 1);
      }
      float ds[32];
      umma::tmem_ld32(t_lane + 32u * ch, ds);
      umma::tmem_ld_wait();
      const float decay = s_ecs[TQ - 1];
#pragma unroll
      for (int j = 0; j < 32; ++j) Grun[j] = decay * Grun[j] + ds[j];
      umma::tc_fence_before(); __syncthreads();
    }
    {   // gradient w.r.t. the state leaving chunk 0
      __nv_bfloat16* gg = p.gstates + ((((long long)db * H + h) * nc) * TN + row) * TP + 32 * ch;
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(gg + 8 * k) = pack8(Grun + 8 * k);
    }
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

// ---- 2. dx / ddt / dA / dD --------------------------------------------------------------------------
constexpr int D2_OFF_C = 0, D2_OFF_B = 2 * HALF, D2_OFF_X = 4 * HALF, D2_OFF_DY = 5 * HALF, D2_OFF_S = 6 * HALF,
              D2_OFF_G = 7 * HALF, D2_OFF_K = 8 * HALF, D2_OFF_TAB = 10 * HALF;
constexpr int D2_OFF_EXTRA = D2_OFF_TAB + 2 * TAB_BYTES;
constexpr int D2_OFF_BAR = D2_OFF_EXTRA + 2 * EXTRA_FLOATS * 4;     // partial sums double-buffered by item parity
constexpr int D2_SMEM = D2_OFF_BAR + 64 + 1024;

template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
ssd_bwd_dx_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                     const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmG,
                     const BwdParams p) {
  constexpr int NCG = NT / 128, NB = 4 / NCG, NTAB = NT - 32;
  static_assert(NCG == 4, "the balanced score-tile split assumes 16 warps");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + D2_OFF_C; uint8_t* sB = base + D2_OFF_B; uint8_t* sX = base + D2_OFF_X;
  uint8_t* sdY = base + D2_OFF_DY; uint8_t* sS = base + D2_OFF_S; uint8_t* sG = base + D2_OFF_G;
  uint8_t* sK = base + D2_OFF_K;
  float* tabs = reinterpret_cast<float*>(base + D2_OFF_TAB);
  float* xtra = reinterpret_cast<float*>(base + D2_OFF_EXTRA);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(base + D2_OFF_BAR);
  uint64_t* bar1 = bar_load + 1;
  uint64_t* bar2 = bar_load + 2;
  uint64_t* bar1b = bar_load + 3;
  uint64_t* bar_cb = bar_load + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2, row = lq * 32 + lane;
  const bool issuer = tid == NTAB;
  static_assert(NCG <= 4 && NT <= BWD_THREADS, "partial-sum slots");
  if (tid == 0) {
    umma::prefetch_tmap(&tmX); umma::prefetch_tmap(&tmDY); umma::prefetch_tmap(&tmS); umma::prefetch_tmap(&tmG);
    umma::mbar_init(bar_load, 1); umma::mbar_init(bar1, 1); umma::mbar_init(bar2, 1); umma::mbar_init(bar1b, 1);
    umma::mbar_init(bar_cb, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  // A work item = (row, chunk, group of Hh heads): C, B (64 of the 144 KB a head-step used to load) and the score
  // tile G = C B^T do not depend on the head, so they are fetched / computed once per item and G stays in TMEM.
  const int H = p.H, L = p.L, di = p.di, nc = p.nc, Hh = p.dxHh, n_items = p.ndirB * nc * (H / Hh);
  constexpr uint32_t idesc_kk128 = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_km64 = umma::make_idesc_bf16(128, 64, 0, 1);
  constexpr uint32_t idesc_mm64 = umma::make_idesc_bf16(128, 64, 1, 1);
  auto load_cb = [&](int item) {
    int base, hg, db, c; p.dHGx.divmod(item, base, hg); p.chunk_of(base, db, c);
    umma::mbar_expect_tx(bar_cb, 4 * HALF);
    umma::tma_load_3d(sC, &tmX, bar_cb, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_cb, di + TN + 64, c * TQ, db);
    umma::tma_load_3d(sB, &tmX, bar_cb, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_cb, di + 64, c * TQ, db);
  };
  auto load_head = [&](int item, int hh, int buf) {
    int base, hg, db, c; p.dHGx.divmod(item, base, hg); p.chunk_of(base, db, c);
    const int h = hg * Hh + hh;
    const int srow = (((db * H + h) * nc) + c) * TN;
    umma::mbar_expect_tx(bar_load, 4 * HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)(srow / TN) * TAB_FLOATS, TAB_BYTES, bar_load);
    umma::tma_load_3d(sX, &tmX, bar_load, h * TP, c * TQ, db);
    umma::tma_load_3d(sdY, &tmDY, bar_load, h * TP, c * TQ, db);
    umma::tma_load_2d(sS, &tmS, bar_load, 0, srow);
    umma::tma_load_2d(sG, &tmG, bar_load, 0, srow);
  };
  uint32_t iseq = 0, seq = 0;                                          // item / head-step sequence numbers of this CTA
  for (int i = tid; i < 2 * HALF / 16; i += NT) reinterpret_cast<uint4*>(sK)[i] = make_uint4(0, 0, 0, 0);   // finite padding rows
  umma::fence_async_smem();
  if (blockIdx.x < n_items && issuer) { load_cb(blockIdx.x); load_head(blockIdx.x, 0, 0); }
  for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++iseq) {
   int base, hg, db, c; p.dHGx.divmod(it, base, hg); p.chunk_of(base, db, c);
   const int dir = p.dB.div(db);
   const int q0 = c * TQ, qv = min(TQ, L - q0);
   const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;              // row blocks / k-steps that hold valid frames
   const long long row0 = (long long)db * L + q0;
   for (int hh = 0; hh < Hh; ++hh, ++seq) {
    const int h = hg * Hh + hh;
    const uint32_t par = seq & 1;
    const float* tab = tabs + (seq & 1) * TAB_FLOATS;
    const float* s_dt = tab + TQ; const float* s_ecs = tab + 3 * TQ; const float* s_eq = tab + 4 * TQ;
    // partial sums of this head-step, one slot per (column group, row) or per warp (shared-memory float atomics are
    // CAS loops); double-buffered by parity so that warp 0 can finish step i while the others start step i+1
    float* s_dcsA = xtra + (seq & 1) * EXTRA_FLOATS;                  // [4][128]  d cs_t, row terms (epilogue A)
    float* s_dcsB = s_dcsA + 4 * TQ;                                  // [4][128]  d cs_q, column terms (epilogue B)
    float* s_ddtx = s_dcsB + 4 * TQ;                                  // [4][128]  <du_q, x_q>
    // per THREAD (summed by the tail warp: a warp reduction here is a 5-deep shuffle chain on every warp's critical path)
    float* s_pdot = s_ddtx + 4 * TQ;                                  // [NT] <Gst, S_in>
    float* s_psc = s_pdot + NT;                                       // [NT] d cs_last from the chunk state
    float* s_pdd = s_psc + NT;                                        // [NT] dD
    const float A = -__expf(p.A_log[dir * H + h]);
    const float Dh = p.Dskip[dir * H + h];
    int nit = it, nh = hh + 1;                                         // the head-step after this one
    if (nh == Hh) { nit = it + gridDim.x; nh = 0; }
    long long tk0 = HNB_CLK();
    umma::mbar_wait(bar_load, par);
    long long tk1 = HNB_CLK();
    if (issuer) {
      if (hh == 0) umma::mbar_wait(bar_cb, iseq & 1);
      umma::tc_fence_after();
      if (hh == 0) {
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // G = C B^T  (kept in TMEM for the item's other heads)
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          umma::mma_bf16_ss(tmem + 0, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024), idesc_kk128, kb > 0);
        }
      }
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)                                   // R = dY X^T
        umma::mma_bf16_ss(tmem + 128, umma::make_smem_desc(umma::smem_u32(sdY) + kb * 32, 16, 1024),
                          umma::make_smem_desc(umma::smem_u32(sX) + kb * 32, 16, 1024), idesc_kk128, kb > 0);
      umma::mma_commit(bar1);                                          // epilogue A starts on G and R ...
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {                                 // Yo = C S_in (unscaled)
        const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
        umma::mma_bf16_ss(tmem + 256, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                          umma::make_smem_desc(umma::smem_u32(sS) + kb * 2048, 1024, 1024), idesc_km64, kb > 0);
      }
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {                                 // du2 = B Gst (independent of epilogue A)
        const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
        umma::mma_bf16_ss(tmem + 384, umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024),
                          umma::make_smem_desc(umma::smem_u32(sG) + kb * 2048, 1024, 1024), idesc_km64, kb > 0);
      }
      umma::mma_commit(bar1b);                                         // ... while these two still run
    }
    // while the MMAs run: the decay term e^{cs_last} <Gst, S_in>
    {
      float dot = 0.f;
      for (int i = tid; i < TQ * 8; i += NT) {                         // same swizzle on both tiles: elementwise product
        float a[8], b[8];
        unpack8(reinterpret_cast<const uint4*>(sS)[i], a);
        unpack8(reinterpret_cast<const uint4*>(sG)[i], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) dot += a[k] * b[k];
      }
      s_pdot[tid] = dot;                                               // scaled by e^{cs_last} in the tail
    }
    umma::mbar_wait(bar1, par);
    umma::tc_fence_after();
    long long tk2 = HNB_CLK();
    // ---- epilogue A (thread = row t, 128/NCG of the 128 columns): K = G o L -> smem;
    //      d cs_t += sum_q W G  +  e^{cs_t} <dY_t, Yo_t>
    if ((row >> 5) >= nblk) {
      s_dcsA[cg * TQ + row] = 0.f;                                     // padding rows: K rows stay finite, see the k-trimmed MMA
    } else {
      const int t = row, I = t >> 5;
      float acc = 0.f;
      {
        const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
        float g[4][8], r[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k <= I) {
            umma::tmem_ld8(t_lane + (uint32_t)(32 * k + 8 * cg), g[k]);
            umma::tmem_ld8(t_lane + 128u + (uint32_t)(32 * k + 8 * cg), r[k]);
          }
        umma::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s0 = 32 * k + 8 * cg;
          if (k <= I) {
            float l[8], d8[8];
            decay8(l, t, I, k, s0, cs_t, e_ref, tab);
            load8(d8, s_dt + s0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              g[k][j] *= l[j];
              // the row sums must use K exactly as the tensor core will see it (bf16): the column sums come out of
              // du1 = K^T dY, and the two cancel in the cumulative sum -- any rounding asymmetry would survive
              acc += r[k][j] * d8[j] * __bfloat162float(__float2bfloat16_rn(g[k][j]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[k][j] = 0.f;
          }
          *reinterpret_cast<uint4*>(sK + (k >> 1) * HALF + swz(t, 4 * (k & 1) + cg)) = pack8(g[k]);
        }
      }
      float yd = 0.f;
      umma::mbar_wait(bar1b, par);
      umma::tc_fence_after();
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const int c16 = NB * cg + bb;
        float yo[16];
        umma::tmem_ld16(t_lane + 256u + 16u * c16, yo);
        umma::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float d[8];
          unpack8(*reinterpret_cast<const uint4*>(sdY + swz(t, 2 * c16 + k)), d);
#pragma unroll
          for (int e = 0; e < 8; ++e) yd += d[e] * yo[8 * k + e];
        }
      }
      s_dcsA[cg * TQ + t] = acc + s_ecs[t] * yd;
    }
    umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads();
    long long tk3 = HNB_CLK();
    if (issuer) {
      umma::tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 8; ++kb)                                   // du1 = K^T dY   (k = time: valid frames only)
        if (kb < nkb)
          umma::mma_bf16_ss(tmem + 320, umma::make_smem_desc(umma::smem_u32(sK) + kb * 2048, HALF, 1024),
                            umma::make_smem_desc(umma::smem_u32(sdY) + kb * 2048, 1024, 1024), idesc_mm64, kb > 0);
      umma::mma_commit(bar2);
    }
    // operands of epilogue B that live in TMA-owned buffers: fetch them now so the buffers can be refilled
    const int q = row;
    float xr[16 * NB], dr[16 * NB];
#pragma unroll
    for (int k = 0; k < 2 * NB; ++k) {
      unpack8(*reinterpret_cast<const uint4*>(sX + swz(q, 2 * NB * cg + k)), xr + 8 * k);
      unpack8(*reinterpret_cast<const uint4*>(sdY + swz(q, 2 * NB * cg + k)), dr + 8 * k);
    }
    umma::mbar_wait(bar2, par);
    umma::tc_fence_after();
    long long tk4 = HNB_CLK();
    __syncthreads();                                                   // every smem operand has been consumed
    if (issuer && nit < n_items) {
      if (nh == 0) load_cb(nit);
      load_head(nit, nh, (seq & 1) ^ 1);
    }
    // ---- epilogue B (thread = row q, 64/NCG of the 64 columns): dx, and the remaining d cs terms
    if ((q >> 5) >= nblk) {
      s_dcsB[cg * TQ + q] = 0.f; s_ddtx[cg * TQ + q] = 0.f;
      s_psc[tid] = 0.f; s_pdd[tid] = 0.f;
    } else {
      const float eq = s_eq[q], dtq = s_dt[q];
      float col = 0.f, sc = 0.f, dux = 0.f, dd = 0.f;
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const int c16 = NB * cg + bb;
        float d1[16], d2[16], o[16];
        umma::tmem_ld16(t_lane + 320u + 16u * c16, d1);
        umma::tmem_ld16(t_lane + 384u + 16u * c16, d2);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float xv = xr[16 * bb + j], dv = dr[16 * bb + j];
          const float du = d1[j] + eq * d2[j];
          o[j] = dtq * du + Dh * dv;
          col += d1[j] * xv;
          sc += d2[j] * xv;
          dux += du * xv;
          dd += dv * xv;
        }
        if (q < qv) {
          __nv_bfloat16* og = p.dxc + (row0 + q) * di + h * TP + 16 * c16;
          *reinterpret_cast<uint4*>(og) = pack8(o);
          *reinterpret_cast<uint4*>(og + 8) = pack8(o + 8);
        }
      }
      col *= dtq; sc *= dtq * eq;
      s_dcsB[cg * TQ + q] = -(col + sc);
      s_ddtx[cg * TQ + q] = dux;
      s_psc[tid] = sc; s_pdd[tid] = dd;
    }
    umma::tc_fence_before(); __syncthreads();                          // partial sums complete; du1 / du2 consumed
    // ---- reverse inclusive cumsum of d cs over the chunk -> ddt, dA_log: ONE warp (lane l owns frames 4l..4l+3);
    //      every other warp goes on to the next item
    if (warp == 4 * (NCG - 1)) {                                       // the warp whose epilogue-A block is all zeros
      float v[4] = {0.f, 0.f, 0.f, 0.f}, ddx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < NCG; ++g) {
        const float4 a = *reinterpret_cast<const float4*>(s_dcsA + g * TQ + 4 * lane);
        const float4 b = *reinterpret_cast<const float4*>(s_dcsB + g * TQ + 4 * lane);
        const float4 d = *reinterpret_cast<const float4*>(s_ddtx + g * TQ + 4 * lane);
        v[0] += a.x + b.x; v[1] += a.y + b.y; v[2] += a.z + b.z; v[3] += a.w + b.w;
        ddx[0] += d.x; ddx[1] += d.y; ddx[2] += d.z; ddx[3] += d.w;
      }
      float extra = 0.f, dd = 0.f;
      {
        const float ecl = s_ecs[TQ - 1];
#pragma unroll
        for (int k = 0; k < NT / 128; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(s_pdot + 128 * k + 4 * lane);
          const float4 b = *reinterpret_cast<const float4*>(s_psc + 128 * k + 4 * lane);
          const float4 c4 = *reinterpret_cast<const float4*>(s_pdd + 128 * k + 4 * lane);
          extra += ecl * (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w);
          dd += c4.x + c4.y + c4.z + c4.w;
        }
      }
      extra = warp_sum(extra); dd = warp_sum(dd);
      if (lane == 31) v[3] += extra;                                   // d cs of the chunk's last frame
      v[2] += v[3]; v[1] += v[2]; v[0] += v[1];                        // suffix sums inside the lane
      float suf = v[0];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += u; }
      const float later = suf - v[0];                                  // everything after this lane's frames
      const float4 dt4 = *reinterpret_cast<const float4*>(s_dt + 4 * lane);
      const float dtv[4] = {dt4.x, dt4.y, dt4.z, dt4.w};
      float accA = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int t = 4 * lane + k;
        const float sv = v[k] + later;
        accA += sv * dtv[k];
        if (t < qv) p.ddt[(row0 + t) * H + h] = sv * A + ddx[k];
      }
      accA = warp_sum(accA);
      if (lane == 0) { atomicAdd(p.dA_log + dir * H + h, accA * A); atomicAdd(p.dD + dir * H + h, dd); }
    }
    if (DBG && p.dbg && blockIdx.x == 0 && tid == 160) {                      // a warp that does not run the tail
      const long long tk5 = HNB_CLK();
      p.dbg[0] += tk1 - tk0; p.dbg[1] += tk2 - tk1; p.dbg[2] += tk3 - tk2; p.dbg[3] += tk4 - tk3;
      p.dbg[4] += tk5 - tk4; p.dbg[6] += 1;
    }
   }
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ---- 3. dB / dC ---------------------------------------------------------------------------------------
constexpr int D3_OFF_C = 0, D3_OFF_B = 2 * HALF, D3_OFF_X = 4 * HALF, D3_OFF_DY = 5 * HALF, D3_OFF_XW = 6 * HALF,
              D3_OFF_DYS = 7 * HALF, D3_OFF_S = 8 * HALF, D3_OFF_G = 9 * HALF, D3_OFF_W = 10 * HALF, D3_OFF_TAB = 12 * HALF;
constexpr int D3_OFF_BAR = D3_OFF_TAB + 2 * TAB_BYTES;
constexpr int D3_SMEM = D3_OFF_BAR + 64 + 1024;

template <int NT, bool DBG>
__global__ void __launch_bounds__(NT, 1)
ssd_bwd_dbc_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                      const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmG,
                      const BwdParams p) {
  constexpr int NCG = NT / 128, NB = 4 / NCG, NTAB = NT - 32;
  static_assert(NCG == 4, "the balanced score-tile split assumes 16 warps");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + D3_OFF_C; uint8_t* sB = base + D3_OFF_B; uint8_t* sX = base + D3_OFF_X;
  uint8_t* sdY = base + D3_OFF_DY; uint8_t* sXw = base + D3_OFF_XW; uint8_t* sdYs = base + D3_OFF_DYS;
  uint8_t* sS = base + D3_OFF_S; uint8_t* sG = base + D3_OFF_G; uint8_t* sW = base + D3_OFF_W;
  float* tabs = reinterpret_cast<float*>(base + D3_OFF_TAB);
  uint64_t* bar_cb = reinterpret_cast<uint64_t*>(base + D3_OFF_BAR);
  uint64_t* bar_xd = bar_cb + 1;
  uint64_t* bar_r = bar_cb + 2;
  uint64_t* bar_m = bar_cb + 3;
  uint64_t* bar_sg = bar_cb + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_cb + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2, row = lq * 32 + lane;
  const bool issuer = tid == NTAB;
  if (tid == 0) {
    umma::prefetch_tmap(&tmX); umma::prefetch_tmap(&tmDY); umma::prefetch_tmap(&tmS); umma::prefetch_tmap(&tmG);
    umma::mbar_init(bar_cb, 1); umma::mbar_init(bar_xd, 1); umma::mbar_init(bar_r, 1); umma::mbar_init(bar_m, 1);
    umma::mbar_init(bar_sg, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  // an item = (row, chunk, head group): with all heads in one CTA the 320 / 160 items of the outer / main stack fill
  // 2.2 / 1.1 waves of 148 SMs; head groups (partial dB/dC sums, added up by the conv backward) fill whole waves
  const int H = p.H, L = p.L, di = p.di, nc = p.nc, HG = p.HG, Hh = H / HG, n_items = p.ndirB * nc * HG;
  constexpr uint32_t i_kk = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t i_km = umma::make_idesc_bf16(128, 128, 0, 1);
  constexpr uint32_t i_mm = umma::make_idesc_bf16(128, 128, 1, 1);
  auto load_cb = [&](int item) {
    int db, c; p.chunk_of(p.dHG.div(item), db, c);
    umma::mbar_expect_tx(bar_cb, 4 * HALF);
    umma::tma_load_3d(sC, &tmX, bar_cb, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_cb, di + TN + 64, c * TQ, db);
    umma::tma_load_3d(sB, &tmX, bar_cb, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_cb, di + 64, c * TQ, db);
  };
  // Per head-step two load groups: X | dY | tables (free again as soon as R and the scaled copies are made) and
  // S_in | Gst (free when the second MMA group retires), so the first group of the NEXT step is in flight, and its
  // R = dY X^T already queued on the tensor pipe, while this step's second MMA group runs.
  auto load_xd = [&](int item, int hh, int buf) {               // hh: head index inside the item's group
    int base, hg, db, c; p.dHG.divmod(item, base, hg); p.chunk_of(base, db, c);
    const int h = hg * Hh + hh;
    const int srow = ((db * H + h) * nc) + c;
    umma::mbar_expect_tx(bar_xd, 2 * HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)srow * TAB_FLOATS, TAB_BYTES, bar_xd);
    umma::tma_load_3d(sX, &tmX, bar_xd, h * TP, c * TQ, db);
    umma::tma_load_3d(sdY, &tmDY, bar_xd, h * TP, c * TQ, db);
  };
  auto load_sg = [&](int item, int hh) {
    int base, hg, db, c; p.dHG.divmod(item, base, hg); p.chunk_of(base, db, c);
    const int h = hg * Hh + hh;
    const int srow = (((db * H + h) * nc) + c) * TN;
    umma::mbar_expect_tx(bar_sg, 2 * HALF);
    umma::tma_load_2d(sS, &tmS, bar_sg, 0, srow);
    umma::tma_load_2d(sG, &tmG, bar_sg, 0, srow);
  };
  auto issue_r = [&]() {                                                // R = dY X^T -> TMEM columns 0..127
    umma::tc_fence_after();
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
      umma::mma_bf16_ss(tmem + 0, umma::make_smem_desc(umma::smem_u32(sdY) + kb * 32, 16, 1024),
                        umma::make_smem_desc(umma::smem_u32(sX) + kb * 32, 16, 1024), i_kk, kb > 0);
    umma::mma_commit(bar_r);
  };
  uint32_t iseq = 0, hseq = 0;
  for (int i = tid; i < 2 * HALF / 16; i += NT) reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);   // finite padding rows
  umma::fence_async_smem();
  if (blockIdx.x < n_items && issuer) {
    load_cb(blockIdx.x); load_xd(blockIdx.x, 0, 0); load_sg(blockIdx.x, 0);
    umma::mbar_wait(bar_xd, 0);
    issue_r();
  }
  for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++iseq) {
    int base, hg, db, c; p.dHG.divmod(it, base, hg); p.chunk_of(base, db, c);
    const int q0 = c * TQ, qv = min(TQ, L - q0);
    const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;             // row blocks / k-steps that hold valid frames
    const long long row0 = (long long)db * L + q0;
    for (int h = 0; h < Hh; ++h, ++hseq) {                             // h: head index inside this item's group
      const uint32_t par = hseq & 1;
      const float* tab = tabs + (hseq & 1) * TAB_FLOATS;
      const float* s_dt = tab + TQ; const float* s_w = tab + 2 * TQ; const float* s_ecs = tab + 3 * TQ;
      int nit = it, nh = h + 1;                                        // the head-step after this one
      if (nh == Hh) { nit = it + gridDim.x; nh = 0; }
      const long long tk0 = HNB_CLK();
      umma::mbar_wait(bar_xd, par);
      const long long tk1 = HNB_CLK();
      for (int i = tid; i < TQ * 8; i += NT) {                         // Xw = w_q X,  dYs = e^{cs_t} dY
        const float w = s_w[i >> 3], e = s_ecs[i >> 3];
        float a[8], b[8];
        unpack8(reinterpret_cast<const uint4*>(sX)[i], a);
        unpack8(reinterpret_cast<const uint4*>(sdY)[i], b);
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] *= w; b[k] *= e; }
        reinterpret_cast<uint4*>(sXw)[i] = pack8(a);
        reinterpret_cast<uint4*>(sdYs)[i] = pack8(b);
      }
      umma::mbar_wait(bar_r, par);
      umma::tc_fence_after();
      const long long tk2 = HNB_CLK();
      if ((row >> 5) < nblk) {            // W[t,q] = R[t,q] L[t,q] dt_q; padding rows are never read (k-trimmed MMAs)
        const int t = row, I = t >> 5;
        const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
        float r[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k <= I) umma::tmem_ld8(t_lane + (uint32_t)(32 * k + 8 * cg), r[k]);
        umma::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s0 = 32 * k + 8 * cg;
          if (k <= I) {
            float l[8], d8[8];
            decay8(l, t, I, k, s0, cs_t, e_ref, tab);
            load8(d8, s_dt + s0);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[k][j] *= l[j] * d8[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[k][j] = 0.f;
          }
          *reinterpret_cast<uint4*>(sW + (k >> 1) * HALF + swz(t, 4 * (k & 1) + cg)) = pack8(r[k]);
        }
      }
      umma::fence_async_smem(); umma::tc_fence_before(); __syncthreads();   // W, Xw, dYs written; R, X, dY, tables consumed
      const long long tk3 = HNB_CLK();
      if (issuer) {
        // next step's first load group goes out BEFORE this step's MMAs: the issuing thread stalls on the tensor-core
        // queue while it feeds 24 MMAs, and a TMA request placed behind them would start ~1.5k cycles late
        if (nit < n_items) load_xd(nit, nh, (hseq & 1) ^ 1);
        if (h == 0) umma::mbar_wait(bar_cb, iseq & 1);
        umma::mbar_wait(bar_sg, par);
        umma::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {                               // dC += W B        (k = time q: valid frames only)
          const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
          if (kb < nkb)
            umma::mma_bf16_ss(tmem + 128, umma::make_smem_desc(umma::smem_u32(sW) + o, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024), i_km, (h > 0 || kb > 0));
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)                                 // dC += (e^{cs} dY) S_in^T
          umma::mma_bf16_ss(tmem + 128, umma::make_smem_desc(umma::smem_u32(sdYs) + kb * 32, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sS) + kb * 32, 16, 1024), i_kk, 1u);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)                                 // dB += W^T C      (k = time t: valid frames only)
          if (kb < nkb)
            umma::mma_bf16_ss(tmem + 256, umma::make_smem_desc(umma::smem_u32(sW) + kb * 2048, HALF, 1024),
                              umma::make_smem_desc(umma::smem_u32(sC) + kb * 2048, HALF, 1024), i_mm, (h > 0 || kb > 0));
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)                                 // dB += (w X) Gst^T
          umma::mma_bf16_ss(tmem + 256, umma::make_smem_desc(umma::smem_u32(sXw) + kb * 32, 16, 1024),
                            umma::make_smem_desc(umma::smem_u32(sG) + kb * 32, 16, 1024), i_kk, 1u);
        umma::mma_commit(bar_m);
        if (nit < n_items) {                                           // next step's R, queued behind this step's MMAs
          umma::mbar_wait(bar_xd, par ^ 1);
          issue_r();
        }
      }
      umma::mbar_wait(bar_m, par);
      umma::tc_fence_after();
      if (issuer && nit < n_items) {                                   // S, G (and C, B after the last head) are free
        if (nh == 0) load_cb(nit);
        load_sg(nit, nh);
      }
      if (DBG && p.dbg && blockIdx.x == 0 && tid == 0) {
        const long long tk4 = HNB_CLK();
        p.dbg[8] += tk1 - tk0; p.dbg[9] += tk2 - tk1; p.dbg[10] += tk3 - tk2; p.dbg[11] += tk4 - tk3; p.dbg[14] += 1;
      }
    }
    // ---- write dB | dC of this chunk (bf16, like the rest of the activation gradients)
    {
      const int t = row;
#pragma unroll
      for (int part = 0; part < 2; ++part) {                            // 0: dB (TMEM 256..383), 1: dC (TMEM 128..255)
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          const int J = NB * cg + bb;
          float v[32];
          umma::tmem_ld32(t_lane + (part == 0 ? 256u : 128u) + 32u * J, v);
          umma::tmem_ld_wait();
          if (t < qv) {
            __nv_bfloat16* og = p.dBC + hg * p.dbc_part_stride + (row0 + t) * (2 * TN) + part * TN + 32 * J;
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(og + 8 * k) = pack8(v + 8 * k);
          }
        }
      }
    }
    umma::tc_fence_before(); __syncthreads();
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}


// ---- 2 + 3 fused: dx / ddt / dA / dD AND dB / dC in ONE pass over (row, chunk [, head group]) ---------------------
// The two kernels above read x | B | C, dY, the chunk states, the state gradients and the tables twice, compute
// R = dY X^T twice and the decay matrix L twice (498 MB of DRAM traffic for 147 MB of algorithmic bytes, two latency
// chains per head).  Here a work item loads C | B once, then per head X, dY, S_in, Gst and the tables once:
//     G = C B^T, R = dY X^T                               (TMEM 256..383, 384..511)
//     epilogue A: one pass over the lower triangle builds BOTH score tiles from one decay evaluation:
//         K = G o L  -> smem (bf16),   W = R o L o dt -> smem (bf16),   row sums of W o G
//         dYs = e^{cs} dY, Xw = w X  -> TMEM as bf16 A operands (448..479, 480..511): no scaled copies in shared memory
//     du1 = K^T dY, du2 = B Gst, Yo = C S_in              (TMEM 256..319, 320..383, 384..447: G and R are dead)
//     epilogue B (dx, ddt terms) takes these into registers, and only THEN
//     dC += W B + dYs S_in^T,  dB += W^T C + Xw Gst^T     (TMEM 0..127, 128..255, accumulated over the item's heads)
// are put on the tensor pipe: tcgen05.ld / tcgen05.st queue behind in-flight MMAs (measured: with the accumulation issued
// early, epilogue B's three TMEM loads took 2.5 k cycles), so the accumulation runs while epilogue B computes and while
// the next head-step's G, R are being set up.  G is recomputed per head (8 MMAs on an otherwise idle pipe) so that its
// TMEM columns can be reused.
// Who issues: an issuing thread stalls on the tensor queue (60 MMAs = ~3.5 k cycles per head-step).  In the first
// version of this kernel that stall sat on everybody's critical path (15.3 k cycles per head-step: no faster than the two
// kernels it replaces); a 17th, issue-only warp rounds the CTA up to 20 warps' worth of registers (96 per thread, 840 B
// of spills).  So two lanes of two LIGHT compute warps (TMEM lane quarter 0: one 32 x 32 block of the triangular score
// tile instead of four) issue, each at a point where its warp would be waiting for the tensor pipe anyway: lane 0 of
// warp 8 issues every TMA copy, G | R and the du / Yo group; lane 0 of warp 4 issues the accumulation group.  Measured
// alternatives (cycles per full head-step at the outer-stack shape; this split: 10.2 k): one lane issuing everything
// 13.5-14.2 k (its warp's own epilogues end up in series with 60 MMAs); a third lane for G | R 11.8 k; <Gst, S_in> by the
// light warps only 12.1 k (the issuing warps ARE light warps: extra work on them delays the next MMA group); x / dY rows
// kept in registers from epilogue A so that the tiles go back to TMA earlier 11.7 k; x / dY re-read from global memory
// for epilogue B 13.5 k.  A head-step of a nearly empty last chunk (L = 398: 14 frames) still costs ~10 k cycles: the step
// is a chain of five mbarrier hand-offs and three MMA groups whose latencies do not shrink with the frame count.
// Shared memory: 12 tiles (192 KB) + tables + partial sums.
constexpr int DF_OFF_C = 0, DF_OFF_B = 2 * HALF, DF_OFF_X = 4 * HALF, DF_OFF_DY = 5 * HALF, DF_OFF_S = 6 * HALF,
              DF_OFF_G = 7 * HALF, DF_OFF_K = 8 * HALF, DF_OFF_W = 10 * HALF, DF_OFF_TAB = 12 * HALF;
constexpr int DF_OFF_EXTRA = DF_OFF_TAB + 2 * TAB_BYTES;
constexpr int DF_OFF_BAR = DF_OFF_EXTRA + 2 * EXTRA_FLOATS * 4;
constexpr int DF_SMEM = DF_OFF_BAR + 128 + 1024;
constexpr int DF_THREADS = BWD_THREADS;
static_assert(DF_SMEM <= 232448, "fused SSD backward: shared memory");
static_assert(DF_OFF_EXTRA % 16 == 0 && DF_OFF_BAR % 8 == 0, "fused SSD backward: alignment");

template <bool DBG>
__global__ void __launch_bounds__(DF_THREADS, 1)
ssd_bwd_fused_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                        const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmG,
                        const BwdParams p, const SpanCuts cuts) {
  pdl_trigger();
  constexpr int NT = BWD_THREADS, NCG = NT / 128;
  static_assert(NCG == 4, "the balanced score-tile split assumes 16 compute warps");
  constexpr uint32_t TM_DC = 0, TM_DB = 128, TM_G = 256, TM_R = 384, TM_DU1 = 256, TM_DU2 = 320, TM_YO = 384,
                     TM_DYS = 448, TM_XW = 480;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + DF_OFF_C; uint8_t* sB = base + DF_OFF_B; uint8_t* sX = base + DF_OFF_X;
  uint8_t* sdY = base + DF_OFF_DY; uint8_t* sS = base + DF_OFF_S; uint8_t* sG = base + DF_OFF_G;
  uint8_t* sK = base + DF_OFF_K; uint8_t* sW = base + DF_OFF_W;
  float* tabs = reinterpret_cast<float*>(base + DF_OFF_TAB);
  float* xtra = reinterpret_cast<float*>(base + DF_OFF_EXTRA);
  uint64_t* bar_cb = reinterpret_cast<uint64_t*>(base + DF_OFF_BAR);
  uint64_t* bar_x = bar_cb + 1;     // TMA: X of the head-step has landed
  uint64_t* bar_dy = bar_cb + 2;    // TMA: dY + tables
  uint64_t* bar_sg = bar_cb + 3;    // TMA: S_in + Gst
  uint64_t* bar_gr = bar_cb + 4;    // MMA: G, R retired
  uint64_t* bar_du1 = bar_cb + 5;   // MMA: du1 retired (dY tile free)
  uint64_t* bar_du = bar_cb + 6;    // MMA: du1, du2, Yo retired
  uint64_t* bar_acc = bar_cb + 7;   // MMA: dC / dB accumulation of the head-step retired
  uint64_t* bar_kw = bar_cb + 8;    // 16 compute warps: K, W, dYs, Xw written; X tile and G, R consumed
  uint64_t* bar_ldb = bar_cb + 9;   // 16 compute warps: epilogue B holds du1, du2, Yo in registers
  uint64_t* bar_out = bar_cb + 10;  // 16 compute warps: dB | dC of the item have left TMEM
  uint64_t* bar_2b = bar_cb + 11;   // the accumulation group is on the tensor pipe: the next G | R may be queued behind it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_cb + 12);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = (warp >> 2) & 3, row = lq * 32 + lane;
  if (tid == 0) {
    umma::prefetch_tmap(&tmX); umma::prefetch_tmap(&tmDY); umma::prefetch_tmap(&tmS); umma::prefetch_tmap(&tmG);
#pragma unroll
    for (int i = 0; i < 8; ++i) umma::mbar_init(bar_cb + i, 1);
    umma::mbar_init(bar_kw, NT / 32); umma::mbar_init(bar_ldb, NT / 32); umma::mbar_init(bar_out, NT / 32);
    umma::mbar_init(bar_2b, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 4 * HALF / 16; i += DF_THREADS) reinterpret_cast<uint4*>(sK)[i] = make_uint4(0, 0, 0, 0);   // sK | sW: finite padding rows
  umma::fence_async_smem();
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  const int H = p.H, L = p.L, di = p.di, nc = p.nc;
  // this CTA's work: piece j = heads [hb, he) of item `bs`
  const long long t_enter = HNB_CLK();
  const int g0 = cuts.g[blockIdx.x], g1 = cuts.g[blockIdx.x + 1];
  const int b0 = p.dH.div(g0);
  auto seg_at = [&](int j, int& bs, int& hb, int& he) -> bool {
    const int g = j == 0 ? g0 : (b0 + j) * H;
    if (g >= g1) return false;
    bs = b0 + j; hb = g - bs * H;
    he = min(g1 - bs * H, H);
    return true;
  };

  const bool isA = tid == 8 * 32;                                      // TMA, G | R, du1 | du2 | Yo
  const bool isB = tid == 4 * 32;                                      // dC | dB accumulation
  constexpr uint32_t i_kk128 = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t i_km128 = umma::make_idesc_bf16(128, 128, 0, 1);
  constexpr uint32_t i_mm128 = umma::make_idesc_bf16(128, 128, 1, 1);
  constexpr uint32_t i_km64 = umma::make_idesc_bf16(128, 64, 0, 1);
  constexpr uint32_t i_mm64 = umma::make_idesc_bf16(128, 64, 1, 1);
  auto load_cb = [&](int bs) {
    int db, c; p.chunk_of(bs, db, c);
    umma::mbar_expect_tx(bar_cb, 4 * HALF);
    umma::tma_load_3d(sC, &tmX, bar_cb, di + TN, c * TQ, db);
    umma::tma_load_3d(sC + HALF, &tmX, bar_cb, di + TN + 64, c * TQ, db);
    umma::tma_load_3d(sB, &tmX, bar_cb, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_cb, di + 64, c * TQ, db);
  };
  auto load_x = [&](int bs, int h) {
    int db, c; p.chunk_of(bs, db, c);
    umma::mbar_expect_tx(bar_x, HALF);
    umma::tma_load_3d(sX, &tmX, bar_x, h * TP, c * TQ, db);
  };
  auto load_dy = [&](int bs, int h, int buf) {
    int db, c; p.chunk_of(bs, db, c);
    const int srow = ((db * H + h) * nc) + c;
    umma::mbar_expect_tx(bar_dy, HALF + TAB_BYTES);
    umma::bulk_load(tabs + buf * TAB_FLOATS, p.tables + (long long)srow * TAB_FLOATS, TAB_BYTES, bar_dy);
    umma::tma_load_3d(sdY, &tmDY, bar_dy, h * TP, c * TQ, db);
  };
  auto load_sg = [&](int bs, int h) {
    int db, c; p.chunk_of(bs, db, c);
    const int srow = (((db * H + h) * nc) + c) * TN;
    umma::mbar_expect_tx(bar_sg, 2 * HALF);
    umma::tma_load_2d(sS, &tmS, bar_sg, 0, srow);
    umma::tma_load_2d(sG, &tmG, bar_sg, 0, srow);
  };
  {
    int bs, hb, he;
    if (isA && seg_at(0, bs, hb, he)) { load_cb(bs); load_x(bs, hb); load_dy(bs, hb, 0); load_sg(bs, hb); }
  }
  {
    uint32_t iseq = 0, seq = 0;                                        // segment / head-step sequence numbers of this CTA
    int bs, hb, he;
    for (int sj = 0; seg_at(sj, bs, hb, he); ++sj, ++iseq) {
      int db, c; p.chunk_of(bs, db, c);
      const int dir = p.dB.div(db);
      const int q0 = c * TQ, qv = min(TQ, L - q0);
      const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;           // row blocks / k-steps that hold valid frames
      const long long row0 = (long long)db * L + q0;
      for (int h = hb; h < he; ++h, ++seq) {
        const uint32_t par = seq & 1;
        int nbs = bs, nh = h + 1;                                      // the head-step after this one
        bool has_next = true;
        if (nh == he) { int e2; has_next = seg_at(sj + 1, nbs, nh, e2); }
        const float* tab = tabs + (seq & 1) * TAB_FLOATS;
        const float* s_dt = tab + TQ; const float* s_w = tab + 2 * TQ; const float* s_ecs = tab + 3 * TQ;
        const float* s_eq = tab + 4 * TQ;
        float* s_dcsA = xtra + (seq & 1) * EXTRA_FLOATS;              // [4][128]  d cs_t, row sums of W o G (epilogue A)
        float* s_dcsB = s_dcsA + 4 * TQ;                              // [4][128]  d cs_q, column and state terms (epilogue B)
        float* s_ddtx = s_dcsB + 4 * TQ;                              // [4][128]  <du_q, x_q>
        float* s_pdot = s_ddtx + 4 * TQ;                              // [NT] <Gst, S_in>   (per thread, summed by the tail warp)
        float* s_psc = s_pdot + NT;                                   // [NT] d cs_last from the chunk state
        float* s_pdd = s_psc + NT;                                    // [NT] dD
        const float A = -__expf(p.A_log[dir * H + h]);
        const float Dh = p.Dskip[dir * H + h];
        const bool tmr = DBG && p.dbg && blockIdx.x == 0 && qv == TQ && (tid == 0 || tid == NT - 31);
        long long* dbg = p.dbg + (tid == 0 ? 0 : 16);
        const long long tk0 = HNB_CLK();
        if (isA) {
          if (h == hb) {
            if (seq > 0) {                                             // C, B, S, Gst are free once the last accumulation retired
              umma::mbar_wait(bar_acc, par ^ 1);
              load_cb(bs); load_sg(bs, hb);
            }
            umma::mbar_wait(bar_cb, iseq & 1);
          }
          umma::mbar_wait(bar_x, par); umma::mbar_wait(bar_dy, par);
          if (seq > 0) umma::mbar_wait(bar_2b, par ^ 1);               // G | R must enter the pipe BEHIND the last accumulation group
          umma::tc_fence_after();
          // ---- group 1: G = C B^T, R = dY X^T.  They overwrite du1 | du2 | Yo (in registers since bar_ldb of the last
          //      head-step) and dYs | Xw (operands of the accumulation issued before them: the pipe runs in order)
  #pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
            umma::mma_bf16_ss(tmem + TM_G, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024), i_kk128, kb > 0);
          }
  #pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            umma::mma_bf16_ss(tmem + TM_R, umma::make_smem_desc(umma::smem_u32(sdY) + kb * 32, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sX) + kb * 32, 16, 1024), i_kk128, kb > 0);
          umma::mma_commit(bar_gr);
          if (h > hb) {                                                // same item: S, Gst of this head once the last accumulation retired
            umma::mbar_wait(bar_acc, par ^ 1);
            load_sg(bs, h);
          }
        }
        umma::mbar_wait(bar_x, par); umma::mbar_wait(bar_dy, par);     // X, dY and the tables are visible to this thread
        if (seq > 0) umma::mbar_wait(bar_acc, par ^ 1);                // the previous head-step no longer reads sK / sW
        umma::mbar_wait(bar_gr, par);
        umma::tc_fence_after();
        const long long tk1 = HNB_CLK();
        // ---- epilogue A (thread = row t, 8 of every 32 columns): K and W from ONE decay evaluation
        const bool live = (row >> 5) < nblk;                           // (warp-uniform) padding rows: sK / sW rows stay finite (zeroed once)
        if (!live) {
          s_dcsA[cg * TQ + row] = 0.f;
        } else {
          const int t = row, I = t >> 5;
          float acc = 0.f;
          const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
          // two passes of two column blocks: 32 instead of 64 live accumulator values per thread (the kernel is at
          // its 128-register ceiling; spills cost L2 round trips here -- with 227 KB of shared memory there is hardly any L1)
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            float g[2][8], r[2][8];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              if (2 * kh + kk <= I) {
                umma::tmem_ld8(t_lane + TM_G + (uint32_t)(32 * (2 * kh + kk) + 8 * cg), g[kk]);
                umma::tmem_ld8(t_lane + TM_R + (uint32_t)(32 * (2 * kh + kk) + 8 * cg), r[kk]);
              }
            if (2 * kh <= I) umma::tmem_ld_wait();
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = 2 * kh + kk, s0 = 32 * k + 8 * cg;
              const uint32_t so = kh * HALF + swz(t, 4 * kk + cg);
              if (k <= I) {
                float l[8], d8[8];
                decay8(l, t, I, k, s0, cs_t, e_ref, tab);
                load8(d8, s_dt + s0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  g[kk][j] *= l[j];                                    // K
                  const float rd = r[kk][j] * d8[j];
                  r[kk][j] = rd * l[j];                                // W
                  // the row sums must use K exactly as the tensor core will see it (bf16): the column sums come out of
                  // du1 = K^T dY, and the two cancel in the cumulative sum -- any rounding asymmetry would survive
                  acc += rd * __bfloat162float(__float2bfloat16_rn(g[kk][j]));
                }
                *reinterpret_cast<uint4*>(sK + so) = pack8(g[kk]);
                *reinterpret_cast<uint4*>(sW + so) = pack8(r[kk]);
              } else {
                *reinterpret_cast<uint4*>(sK + so) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(sW + so) = make_uint4(0, 0, 0, 0);
              }
            }
          }
          s_dcsA[cg * TQ + t] = acc;
        }
        // The TMEM copies below overwrite R columns 64 + 8 cg .. +7 and 96 + 8 cg .. +7 of this warp's lane quarter: exactly
        // the strips of column blocks 2 and 3 that THIS warp has just read (or never reads) -- program order suffices.
        // scaled copies as TMEM operands: dYs = e^{cs_t} dY, Xw = w_q X  (bf16 pairs, A operands of the K = 64 MMAs),
        // from this thread's 16 columns of its x and dY rows
        {
          const float ecs_t = s_ecs[row], w_q = s_w[row];
          uint32_t pk[8];
          float v[8];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            unpack8(*reinterpret_cast<const uint4*>(sdY + swz(row, 2 * cg + k)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= ecs_t;
            const uint4 q4 = pack8(v);
            pk[4 * k] = q4.x; pk[4 * k + 1] = q4.y; pk[4 * k + 2] = q4.z; pk[4 * k + 3] = q4.w;
          }
          umma::tmem_st8(t_lane + TM_DYS + (uint32_t)(8 * cg), pk);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            unpack8(*reinterpret_cast<const uint4*>(sX + swz(row, 2 * cg + k)), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= w_q;
            const uint4 q4 = pack8(v);
            pk[4 * k] = q4.x; pk[4 * k + 1] = q4.y; pk[4 * k + 2] = q4.z; pk[4 * k + 3] = q4.w;
          }
          umma::tmem_st8(t_lane + TM_XW + (uint32_t)(8 * cg), pk);
          umma::tmem_st_wait();
        }
        const long long tk2 = HNB_CLK();
        umma::fence_async_smem(); umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_kw);                      // K, W, dYs, Xw written; G, R, X consumed
        if (isA) {                                                     // ---- MMA group 2a: what epilogue B needs
          umma::mbar_wait(bar_kw, par);
          umma::tc_fence_after();
          umma::mbar_wait(bar_sg, par);
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)                               // du1 = K^T dY   (k = time: valid frames only)
            if (kb < nkb)
              umma::mma_bf16_ss(tmem + TM_DU1, umma::make_smem_desc(umma::smem_u32(sK) + kb * 2048, HALF, 1024),
                                umma::make_smem_desc(umma::smem_u32(sdY) + kb * 2048, 1024, 1024), i_mm64, kb > 0);
          umma::mma_commit(bar_du1);
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {                             // du2 = B Gst
            const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
            umma::mma_bf16_ss(tmem + TM_DU2, umma::make_smem_desc(umma::smem_u32(sB) + o, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sG) + kb * 2048, 1024, 1024), i_km64, kb > 0);
          }
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {                             // Yo = C S_in (unscaled)
            const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
            umma::mma_bf16_ss(tmem + TM_YO, umma::make_smem_desc(umma::smem_u32(sC) + o, 16, 1024),
                              umma::make_smem_desc(umma::smem_u32(sS) + kb * 2048, 1024, 1024), i_km64, kb > 0);
          }
          umma::mma_commit(bar_du);
        }
        // while the second MMA group runs: the decay term e^{cs_last} <Gst, S_in>
        umma::mbar_wait(bar_sg, par);
        {
          float dot = 0.f;
          for (int i = tid; i < TQ * 8; i += NT) {                     // same swizzle on both tiles: elementwise product
            float a[8], b[8];
            unpack8(reinterpret_cast<const uint4*>(sS)[i], a);
            unpack8(reinterpret_cast<const uint4*>(sG)[i], b);
#pragma unroll
            for (int k = 0; k < 8; ++k) dot += a[k] * b[k];
          }
          s_pdot[tid] = dot;                                           // scaled by e^{cs_last} in the tail
        }
        const long long tk3 = HNB_CLK();
        umma::mbar_wait(bar_du, par);
        umma::tc_fence_after();
        const long long tk4 = HNB_CLK();
        // ---- epilogue B (thread = row q, 16 of the 64 columns): dx, and the remaining d cs terms
        const int q = row;
        float d1[16], d2[16], yo[16];
        uint4 xr4[2], dr4[2];                                          // this thread's 16 columns of its x and dY rows
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          xr4[k] = *reinterpret_cast<const uint4*>(sX + swz(row, 2 * cg + k));
          dr4[k] = *reinterpret_cast<const uint4*>(sdY + swz(row, 2 * cg + k));
        }
        if (live) {
          umma::tmem_ld16(t_lane + TM_DU1 + 16u * cg, d1);
          umma::tmem_ld16(t_lane + TM_DU2 + 16u * cg, d2);
          umma::tmem_ld16(t_lane + TM_YO + 16u * cg, yo);
          umma::tmem_ld_wait();
        }
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_ldb);                     // the accumulation (and the next G, R) may go on the pipe
        if (isA && has_next) {                                         // X, dY tiles are free: du1 retired, every thread holds its rows
          umma::mbar_wait(bar_ldb, par);
          load_x(nbs, nh); load_dy(nbs, nh, (seq & 1) ^ 1);
        }
        if (isB) {                                                     // ---- MMA group 2b: accumulate dC, dB over the item's heads
          umma::mbar_wait(bar_ldb, par);
          if (h == hb && iseq > 0) umma::mbar_wait(bar_out, (iseq - 1) & 1);   // the last item's dB | dC have left TMEM
          umma::tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {                             // dC += W B        (k = time q: valid frames only)
            const uint32_t o = (kb >> 2) * HALF + (kb & 3) * 32;
            if (kb < nkb)
              umma::mma_bf16_ss(tmem + TM_DC, umma::make_smem_desc(umma::smem_u32(sW) + o, 16, 1024),
                                umma::make_smem_desc(umma::smem_u32(sB) + kb * 2048, HALF, 1024), i_km128, (h > hb || kb > 0));
          }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)                               // dC += (e^{cs} dY) S_in^T
            umma::mma_bf16_ts(tmem + TM_DC, tmem + TM_DYS + 8 * kb,
                              umma::make_smem_desc(umma::smem_u32(sS) + kb * 32, 16, 1024), i_kk128, 1u);
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)                               // dB += W^T C      (k = time t: valid frames only)
            if (kb < nkb)
              umma::mma_bf16_ss(tmem + TM_DB, umma::make_smem_desc(umma::smem_u32(sW) + kb * 2048, HALF, 1024),
                                umma::make_smem_desc(umma::smem_u32(sC) + kb * 2048, HALF, 1024), i_mm128, (h > hb || kb > 0));
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)                               // dB += (w X) Gst^T
            umma::mma_bf16_ts(tmem + TM_DB, tmem + TM_XW + 8 * kb,
                              umma::make_smem_desc(umma::smem_u32(sG) + kb * 32, 16, 1024), i_kk128, 1u);
          umma::mma_commit(bar_acc);
          umma::mbar_arrive(bar_2b);
        }
        const long long tk5 = HNB_CLK();
        if (!live) {
          s_dcsB[cg * TQ + q] = 0.f; s_ddtx[cg * TQ + q] = 0.f;
          s_psc[tid] = 0.f; s_pdd[tid] = 0.f;
        } else {
          const float eq = s_eq[q], dtq = s_dt[q];
          float col = 0.f, sc = 0.f, dux = 0.f, dd = 0.f, yd = 0.f;
          float o[16], xv[16], dv[16];
          unpack8(xr4[0], xv); unpack8(xr4[1], xv + 8);
          unpack8(dr4[0], dv); unpack8(dr4[1], dv + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float du = d1[j] + eq * d2[j];
            o[j] = dtq * du + Dh * dv[j];
            col += d1[j] * xv[j];
            sc += d2[j] * xv[j];
            dux += du * xv[j];
            dd += dv[j] * xv[j];
            yd += dv[j] * yo[j];
          }
          if (q < qv) {
            __nv_bfloat16* og = p.dxc + (row0 + q) * di + h * TP + 16 * cg;
            *reinterpret_cast<uint4*>(og) = pack8(o);
            *reinterpret_cast<uint4*>(og + 8) = pack8(o + 8);
          }
          col *= dtq; sc *= dtq * eq;
          s_dcsB[cg * TQ + q] = s_ecs[q] * yd - (col + sc);            // + e^{cs_t} <dY_t, (C S_in)_t>: the state path's row term
          s_ddtx[cg * TQ + q] = dux;
          s_psc[tid] = sc; s_pdd[tid] = dd;
        }
        const long long tk6 = HNB_CLK();
        // partial sums of the head-step complete: only the tail warp waits for everybody, the others go on
        if (warp == 4 * (NCG - 1)) named_sync(1, NT);
        else asm volatile("bar.arrive 1, %0;" ::"r"(NT) : "memory");
        if (tmr) {
          const long long tk7 = HNB_CLK();
          dbg[0] += tk1 - tk0; dbg[1] += tk2 - tk1; dbg[2] += tk3 - tk2; dbg[3] += tk4 - tk3; dbg[4] += tk5 - tk4;
          dbg[5] += tk6 - tk5; dbg[6] += tk7 - tk6; dbg[9] += 1;
          if (h == hb) { dbg[10] += tk7 - tk0; dbg[11] += 1; }         // first head-step of a piece
        }
        // ---- reverse inclusive cumsum of d cs over the chunk -> ddt, dA_log: ONE warp (lane l owns frames 4l..4l+3);
        //      every other warp goes on to the next head-step
        if (warp == 4 * (NCG - 1)) {
          float v[4] = {0.f, 0.f, 0.f, 0.f}, ddx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int gI = 0; gI < NCG; ++gI) {
            const float4 a = *reinterpret_cast<const float4*>(s_dcsA + gI * TQ + 4 * lane);
            const float4 b = *reinterpret_cast<const float4*>(s_dcsB + gI * TQ + 4 * lane);
            const float4 d = *reinterpret_cast<const float4*>(s_ddtx + gI * TQ + 4 * lane);
            v[0] += a.x + b.x; v[1] += a.y + b.y; v[2] += a.z + b.z; v[3] += a.w + b.w;
            ddx[0] += d.x; ddx[1] += d.y; ddx[2] += d.z; ddx[3] += d.w;
          }
          float extra = 0.f, dd = 0.f;
          {
            const float ecl = s_ecs[TQ - 1];
#pragma unroll
            for (int k = 0; k < NT / 128; ++k) {
              const float4 a = *reinterpret_cast<const float4*>(s_pdot + 128 * k + 4 * lane);
              const float4 b = *reinterpret_cast<const float4*>(s_psc + 128 * k + 4 * lane);
              const float4 c4 = *reinterpret_cast<const float4*>(s_pdd + 128 * k + 4 * lane);
              extra += ecl * (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w);
              dd += c4.x + c4.y + c4.z + c4.w;
            }
          }
          extra = warp_sum(extra); dd = warp_sum(dd);
          if (lane == 31) v[3] += extra;                               // d cs of the chunk's last frame
          v[2] += v[3]; v[1] += v[2]; v[0] += v[1];                    // suffix sums inside the lane
          float suf = v[0];
#pragma unroll
          for (int o2 = 1; o2 < 32; o2 <<= 1) { const float u = __shfl_down_sync(0xffffffffu, suf, o2); if (lane + o2 < 32) suf += u; }
          const float later = suf - v[0];                              // everything after this lane's frames
          const float4 dt4 = *reinterpret_cast<const float4*>(s_dt + 4 * lane);
          const float dtv[4] = {dt4.x, dt4.y, dt4.z, dt4.w};
          float accA = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int tt = 4 * lane + k;
            const float sv = v[k] + later;
            accA += sv * dtv[k];
            if (tt < qv) p.ddt[(row0 + tt) * H + h] = sv * A + ddx[k];
          }
          accA = warp_sum(accA);
          if (lane == 0) { atomicAdd(p.dA_log + dir * H + h, accA * A); atomicAdd(p.dD + dir * H + h, dd); }
        }
      }
      // ---- dB | dC of this piece: the item's last heads (hb > 0) go to the fix-up buffer, its first heads (or all of them)
      //      are rounded and stored (bf16, like the rest of the activation gradients) after the neighbour's part was added
      const long long te0 = HNB_CLK();
      umma::mbar_wait(bar_acc, (seq - 1) & 1);
      umma::tc_fence_after();
      const long long te1 = HNB_CLK();
      {
        const int t = row;
        const bool to_fix = hb > 0, from_fix = he < H;
        // fix[k]: [dB | dC][8 float4 of a thread's 32 columns][column group][row] -- a warp's access is 512 contiguous bytes
        float4* fx = reinterpret_cast<float4*>(p.fix + (size_t)(blockIdx.x + (from_fix ? 1 : 0)) * (2 * NCG * TQ * 32)) + cg * TQ + t;
        if (from_fix) {                                                 // (raised ~a whole piece ago: this does not spin in practice)
          const int* fl = p.fix_flags + blockIdx.x + 1;
          int got;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(fl) : "memory");
            if (got < NT / 32) __nanosleep(256);
          } while (got < NT / 32);
        }
#pragma unroll
        for (int bc = 0; bc < 2; ++bc) {                                // 0: dB, 1: dC
          float v[32];
          umma::tmem_ld32(t_lane + (bc == 0 ? TM_DB : TM_DC) + 32u * cg, v);
          umma::tmem_ld_wait();
          if (DBG && p.dbg && blockIdx.x == 0 && tid == 0 && iseq < 7) p.dbg[400 + 4 * iseq + bc] = HNB_CLK() - te1;
          float4* f4 = fx + bc * (8 * NCG * TQ);
          if (to_fix) {
#pragma unroll
            for (int k = 0; k < 8; ++k) __stcg(f4 + k * (NCG * TQ), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
          } else {
            if (from_fix) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 a = __ldcg(f4 + k * (NCG * TQ));
                v[4 * k] += a.x; v[4 * k + 1] += a.y; v[4 * k + 2] += a.z; v[4 * k + 3] += a.w;
              }
            }
            if (t < qv) {                                               // 64 contiguous bytes per thread: two full 32-byte sectors
              __nv_bfloat16* og = p.dBC + (row0 + t) * (2 * TN) + bc * TN + 32 * cg;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint4 lo = pack8(v + 16 * k), hi = pack8(v + 16 * k + 8);
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(og + 16 * k), "r"(lo.x), "r"(lo.y),
                             "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
              }
            }
          }
        }
        if (to_fix) {                                                   // every warp's share is visible before its count
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(p.fix_flags + blockIdx.x, 1);
        }
        if (DBG && p.dbg && blockIdx.x == 0 && tid == 0) {
          p.dbg[12] += te1 - te0; p.dbg[13] += HNB_CLK() - te1; p.dbg[14] += 1;
          if (iseq < 7) { p.dbg[400 + 4 * iseq + 2] = HNB_CLK() - te1; p.dbg[400 + 4 * iseq + 3] = (to_fix ? 1 : 0) + (from_fix ? 2 : 0) + 4; }
        }
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_out);                       // the next item's accumulation may overwrite dC | dB
    }
  }
  umma::tc_fence_before(); __syncthreads();
  if (DBG && p.dbg && tid == 0) {                                      // per-CTA: cycles from entry to here, head-steps done
    p.dbg[32 + 2 * blockIdx.x] = HNB_CLK() - t_enter; p.dbg[33 + 2 * blockIdx.x] = g1 - g0;
  }
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace hnb

using namespace hnb;

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// workspace of the tcgen05 path: bf16 states [ndir*B, H, nc, 128, 64] followed (128-byte aligned) by the fp32 tables
static size_t tc_tables_offset(int ndir, int B, int L, int H) {
  const size_t st = (size_t)ndir * B * H * cdiv(L, TQ) * TN * TP * sizeof(__nv_bfloat16);
  return (st + 127) / 128 * 128;
}
// the backward's second workspace holds the state gradients where the forward's holds the states, and behind them the
// fix-up buffer of the fused kernel's span schedule instead of the tables
constexpr size_t FIX_BYTES = (size_t)SPAN_MAX_CTAS * 2 * 4 * TQ * 32 * sizeof(float) + (SPAN_MAX_CTAS + 1 + 31) / 32 * 32 * sizeof(int);
long long hnb_ssd_tc_ws_bytes(int ndir, int B, int L, int H) {
  const size_t tabs = (size_t)ndir * B * H * cdiv(L, TQ) * TAB_BYTES + 128;
  return (long long)(tc_tables_offset(ndir, B, L, H) + (tabs > FIX_BYTES ? tabs : FIX_BYTES));
}

int hnb_ssd_fwd_split_tc(const CUtensorMap* tmX, const void* xconv, const float* Dskip, const float* tables, int ndir, int B,
                         int L, int di, int H, void* y, void* states, void* stream);

int hnb_ssd_fwd_tc(const void* xconv, const float* dt, const float* A_log, const float* Dskip, int ndir, int B, int L,
                   int di, int N, int H, void* y, void* states, void* stream, int variant_req) {
  HNB_CHECK_ARG(N == TN && di == H * TP, "ssd_fwd(tcgen05): built for d_state=128, headdim=64");
  const int C = di + 2 * N;
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)C, (uint64_t)L, (uint64_t)ndir * B};
  uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)L * C * 2};
  uint32_t box[3] = {64, TQ, 1};
  int rc = make_tmap_bf16(&tm, xconv, 3, dims, strides, box);
  if (rc) return rc;
  FwdParams p;
  p.dt = dt; p.A_log = A_log; p.Dskip = Dskip; p.xconv = (const __nv_bfloat16*)xconv;
  p.y = (__nv_bfloat16*)y; p.states = (__nv_bfloat16*)states;
  p.ndirB = ndir * B; p.B = B; p.L = L; p.H = H; p.di = di; p.nc = cdiv(L, TQ);
  p.dH = FastDiv(H); p.dB = FastDiv(B); p.dnc = FastDiv(p.nc);
  HNB_CHECK_ARG((long long)ndir * B * H * p.nc * (H > B ? H : B) < (1LL << 31), "ssd_fwd(tcgen05): problem too large");
  float* tables = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(states) + tc_tables_offset(ndir, B, L, H));
  p.tables = tables;
  const int items = ndir * B * H;
  int* counter = reinterpret_cast<int*>(tables + (size_t)items * p.nc * TAB_FLOATS);   // the workspace's spare 128 bytes
  hnb::launch_pdl(ssd_tables_kernel, dim3(items * p.nc), dim3(TQ), 0, (cudaStream_t)stream, dt, A_log, tables, B, L, H, p.nc, counter);
  static const bool dynamic = getenv("HNB_SSD_FWD_DYNAMIC") && atoi(getenv("HNB_SSD_FWD_DYNAMIC")) != 0;
  p.counter = dynamic ? counter : nullptr;
  HNB_LAUNCH_CHECK("ssd_tables");
  p.dbg = nullptr;
  const bool debug = getenv("HNB_SSD_DEBUG") != nullptr;
  // 1 / 2: persistent kernel per (row, head), one / two CTAs per SM; 3: states pass + scan over every chunk at once.  Measured
  // (scratch/ssd_time.py, B200): the split pays once a row has many chunks -- 60 s utterances, 12 chunks: 154 vs 167 us per
  // call -- and loses on the 4-chunk rows of the headline batch (141 vs 123 us), where the persistent kernel's chain is short.
  static const int variant_env = getenv("HNB_SSD_FWD") ? atoi(getenv("HNB_SSD_FWD")) : 0;
  const int variant = variant_req > 0 ? variant_req : variant_env > 0 ? variant_env : (p.nc >= 8 ? 3 : 2);
  if (variant == 3)                                   // states pass + scan over every chunk at once (ssd_fwd_split.cu)
    return hnb_ssd_fwd_split_tc(&tm, xconv, Dskip, tables, ndir, B, L, di, H, y, states, stream);
  if (variant == 2 && !debug) {
    static const int per_sm = getenv("HNB_SSD_FWD2_PER_SM") ? atoi(getenv("HNB_SSD_FWD2_PER_SM")) : 2;   // diagnosis
    const int grid2 = items < per_sm * sm_count() ? items : per_sm * sm_count();
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_fwd2_tc_kernel, F2_SMEM));
    hnb::launch_pdl(ssd_fwd2_tc_kernel, dim3(grid2), dim3(F2_THREADS), F2_SMEM, (cudaStream_t)stream, tm, p);
    HNB_LAUNCH_CHECK("ssd_fwd2_tc");
    return HNB_OK;
  }
  const int grid = items < sm_count() ? items : sm_count();
  if (debug) { cudaMalloc(&p.dbg, 64); cudaMemsetAsync(p.dbg, 0, 64, (cudaStream_t)stream); }
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_fwd_tc_kernel<FWD_THREADS, false>, FWD_SMEM));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_fwd_tc_kernel<FWD_THREADS, true>, FWD_SMEM));
  if (debug) ssd_fwd_tc_kernel<FWD_THREADS, true><<<grid, FWD_THREADS, FWD_SMEM, (cudaStream_t)stream>>>(tm, p);
  else ssd_fwd_tc_kernel<FWD_THREADS, false><<<grid, FWD_THREADS, FWD_SMEM, (cudaStream_t)stream>>>(tm, p);
  HNB_LAUNCH_CHECK("ssd_fwd_tc");
  if (debug) {
    long long h[8];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(h, p.dbg, 64, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const double n = h[6] > 0 ? (double)h[6] : 1.0;
    fprintf(stderr, "[ssd_fwd CTA0] steps %lld | cycles/step: wait load %.0f, G mma wait (+Xw) %.0f, epi1 %.0f, "
            "mma2 wait %.0f, epi2+3 %.0f | next load issued %.0f after step start, lands %.0f later\n", h[6], h[0] / n,
            h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[7] / n, h[5] / n);
  }
  return HNB_OK;
}

// Head groups of the dB/dC kernel (1 or 2).  Items are handed out round-robin in descending cost order, so the
// first CTA carries the largest item of every round: its load is the makespan.  Cost unit: one head-step of a full
// chunk; a partly filled chunk costs its fixed latency plus the trimmed share.  A second part costs the convolution
// backward one extra read of dB | dC (about two head-steps' worth of time).
// variant 0: the fused backward kernel (an item carries ALL per-head work of its heads); 1: the three-kernel backward.
static bool ssd_bwd_legacy_env() {
  static const bool legacy = getenv("HNB_SSD_BWD") && atoi(getenv("HNB_SSD_BWD")) == 3;   // 3 = round-1 three-kernel backward
  return legacy;
}
// Span schedule of the fused backward (BwdParams::span_mode): cost of a head-step in 1/64 of a full-chunk head-step, the
// per-item overhead (C | B load, dB | dC store: ~0.6 head-steps) spread over the item's heads.  Applicable when every CTA's
// piece is longer than an item (H head-steps), so that no item is cut twice.
static bool span_enabled() {
  static const bool off = getenv("HNB_SSD_SPAN") && atoi(getenv("HNB_SSD_SPAN")) == 0;
  return !off;
}
static int span_ctas() {
  static const int force = getenv("HNB_SSD_SPAN_CTAS") ? atoi(getenv("HNB_SSD_SPAN_CTAS")) : 0;   // diagnosis: fewer SMs
  const int n = sm_count() < SPAN_MAX_CTAS ? sm_count() : SPAN_MAX_CTAS;
  return force > 0 && force < n ? force : n;
}
// Cut positions (head-step indices) of the G pieces: equal cost.  A head-step is a chain of hand-offs whose latencies
// hardly shrink with the frame count, so a head-step of the partly filled last chunk weighs w = w0 + (1 - w0) rem / 128 of a
// full one with w0 close to 1 (measured: with w0 = 0.45 the CTAs holding the 14-frame chunks of the 398-frame rows ran
// 43 head-steps against 23 elsewhere and the kernel took 279 us instead of 223).  HNB_SSD_SPAN=0: cuts at item borders only.
static void span_cuts(int ndirB, int L, int H, int G, SpanCuts* out) {
  static const double w0 = getenv("HNB_SSD_SPAN_W0") ? atof(getenv("HNB_SSD_SPAN_W0")) : 0.9;
  const int nc = cdiv(L, TQ), rem = L - (nc - 1) * TQ;
  const long long nfull = (long long)ndirB * (rem == TQ ? nc : nc - 1), npart = (long long)ndirB * nc - nfull;
  const double w = w0 + (1.0 - w0) * rem / TQ;
  const double cfull = (double)nfull * H, total = cfull + (double)npart * H * w;
  const long long steps = (nfull + npart) * H;
  if (!span_enabled()) {                                               // whole items, contiguous runs of equal length
    const long long n = nfull + npart;
    for (int k = 0; k <= SPAN_MAX_CTAS; ++k) out->g[k] = (int)((k >= G ? n : n * k / G) * H);
    return;
  }
  out->g[0] = 0;
  double done = 0.0;                                                   // cost in front of the previous cut
  for (int k = 1; k < G; ++k) {
    const double x = done + (total - done) / (G - k + 1);               // what is left, shared by the CTAs that are left
    long long c = x < cfull ? (long long)(x + 0.5) : (long long)(cfull + (x - cfull) / w + 0.5);
    const long long prev = out->g[k - 1];
    if (c < prev) c = prev;
    if (c > steps) c = steps;
    if (prev % H != 0 && c % H != 0 && prev / H == c / H) c = (c / H + 1) * H;   // never two cuts inside one item
    out->g[k] = (int)c;
    done = c < (long long)cfull ? (double)c : cfull + ((double)c - cfull) * w;
  }
  for (int k = G; k <= SPAN_MAX_CTAS; ++k) out->g[k] = (int)steps;
}

// host-only query (tests): the G + 1 cut positions of the span schedule for a problem, G pieces
extern "C" int hnb_ssd_span_cuts(int ndirB, int L, int H, int G, int* out) {
  HNB_CHECK_ARG(out && ndirB > 0 && L > 0 && H > 0 && G >= 1 && G <= SPAN_MAX_CTAS, "ssd_span_cuts: bad arguments");
  SpanCuts c;
  span_cuts(ndirB, L, H, G, &c);
  for (int k = 0; k <= G; ++k) out[k] = c.g[k];
  return HNB_OK;
}

int hnb_ssd_dbc_parts_tc(int ndir, int B, int L, int H, int variant) {
  if (ssd_bwd_legacy_env()) variant = 1;
  if (variant == 0) return 1;                          // fused kernel: cut items are summed inside the kernel (fix-up buffer)
  const int sms = sm_count(), nc = cdiv(L, TQ), rem = L - (nc - 1) * TQ;
  const int nfull = ndir * B * (rem == TQ ? nc : nc - 1), npart = ndir * B * nc - nfull;
  int best = 1;
  double best_cost = 1e30;
  for (int hg = 1; hg <= 2; ++hg) {
    if (H % hg) continue;
    const double full = H / hg + 0.6, part = (0.45 + 0.55 * rem / TQ) * (H / hg) + 0.6;
    // a second part costs the convolution backward one extra read of dB | dC: ~2 head-steps of the dB/dC kernel alone,
    // ~1 of the (longer) fused head-step
    double cost = (variant == 0 ? 1.0 : 2.0) * (hg - 1);
    for (long long i = 0; i < (long long)(nfull + npart) * hg; i += sms) cost += i < (long long)nfull * hg ? full : part;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = hg; }
  }
  return best;
}

// Heads per work item of the dx kernel.  More heads per item amortise the C | B loads and the G = C B^T product, fewer
// heads balance the static round-robin better; the first CTA carries the largest item of every round (descending cost
// order), so its load is the makespan.  Cost unit: one head-step of a full chunk with nothing shared.
static int dx_heads_per_item(int ndirB, int L, int H, int sms) {
  const int nc = cdiv(L, TQ), rem = L - (nc - 1) * TQ;
  const long long nfull = (long long)ndirB * (rem == TQ ? nc : nc - 1), npart = (long long)ndirB * nc - nfull;
  int best = 1;
  double best_cost = 1e30;
  for (int hh = 1; hh <= 8; ++hh) {
    if (H % hh) continue;
    const int groups = H / hh;
    const double step = 1.0 - 0.12 * (1.0 - 1.0 / hh), full = hh * step, part = (0.45 + 0.55 * rem / TQ) * full;
    double cost = 0.0;
    for (long long i = 0; i < (nfull + npart) * groups; i += sms) cost += i < nfull * groups ? full : part;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = hh; }
  }
  return best;
}

int hnb_ssd_bwd_tc(const void* dy, const void* xconv, const void* y, const float* dt, const float* A_log,
                   const float* Dskip, const void* states, int ndir, int B, int L, int di, int N, int H, void* dxc,
                   void* dBC, int dbc_parts, float* ddt, float* dA_log, float* dD, void* ws2, void* stream, int variant) {
  (void)y;
  if (ssd_bwd_legacy_env()) variant = 1;
  HNB_CHECK_ARG(N == TN && di == H * TP, "ssd_bwd(tcgen05): built for d_state=128, headdim=64");
  const int C = di + 2 * N, nc = cdiv(L, TQ);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmX, tmDY, tmS, tmG;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)L, (uint64_t)ndir * B};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)L * C * 2};
    uint32_t box[3] = {64, TQ, 1};
    if ((rc = make_tmap_bf16(&tmX, xconv, 3, dims, strides, box))) return rc;
    dims[0] = (uint64_t)di; strides[0] = (uint64_t)di * 2; strides[1] = (uint64_t)L * di * 2;
    if ((rc = make_tmap_bf16(&tmDY, dy, 3, dims, strides, box))) return rc;
    uint64_t d2[2] = {(uint64_t)TP, (uint64_t)ndir * B * H * nc * TN};
    uint64_t s2[1] = {(uint64_t)TP * 2};
    uint32_t b2[2] = {TP, TN};
    if ((rc = make_tmap_bf16(&tmS, states, 2, d2, s2, b2))) return rc;
    if ((rc = make_tmap_bf16(&tmG, ws2, 2, d2, s2, b2))) return rc;
  }
  BwdParams p;
  p.dt = dt; p.A_log = A_log; p.Dskip = Dskip; p.gstates = (__nv_bfloat16*)ws2; p.dxc = (__nv_bfloat16*)dxc;
  p.xconv = (const __nv_bfloat16*)xconv; p.dy = (const __nv_bfloat16*)dy;
  p.dBC = (__nv_bfloat16*)dBC; p.ddt = ddt; p.dA_log = dA_log; p.dD = dD;
  HNB_CHECK_ARG((dbc_parts == 1 || dbc_parts == 2) && H % dbc_parts == 0, "ssd_bwd(tcgen05): dbc_parts must be 1 or 2 and divide H");
  p.HG = dbc_parts; p.dbc_part_stride = (long long)ndir * B * L * 2 * N;
  p.ndirB = ndir * B; p.B = B; p.L = L; p.H = H; p.di = di; p.nc = nc;
  p.dH = FastDiv(H); p.dB = FastDiv(B); p.dnc = FastDiv(nc);
  p.nfc = (L % TQ == 0) ? nc : nc - 1; p.n_full = p.ndirB * p.nfc;
  p.dnfc = FastDiv(p.nfc > 0 ? p.nfc : 1); p.dHG = FastDiv(dbc_parts);
  {
    static const int force = getenv("HNB_SSD_DX_HEADS") ? atoi(getenv("HNB_SSD_DX_HEADS")) : 0;   // tuning knob
    p.dxHh = (force > 0 && H % force == 0) ? force : dx_heads_per_item(p.ndirB, L, H, sm_count());
    p.dHGx = FastDiv(H / p.dxHh);
  }
  HNB_CHECK_ARG((long long)ndir * B * H * nc * (H > B ? H : B) < (1LL << 31), "ssd_bwd(tcgen05): problem too large");
  p.tables = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(states) + tc_tables_offset(ndir, B, L, H));
  p.dbg = nullptr;
  HNB_CHECK_ARG(variant != 0 || dbc_parts == 1, "ssd_bwd(tcgen05): the fused backward writes one dBC part");
  p.fix = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws2) + tc_tables_offset(ndir, B, L, H));
  p.fix_flags = reinterpret_cast<int*>(p.fix + (size_t)SPAN_MAX_CTAS * 2 * 4 * TQ * 32);
  SpanCuts cuts = {};
  if (variant == 0) span_cuts(p.ndirB, L, H, span_ctas(), &cuts);
  const bool debug = getenv("HNB_SSD_DEBUG") != nullptr;
  if (debug) { cudaMalloc(&p.dbg, 4096); cudaMemsetAsync(p.dbg, 0, 4096, st); }
  const int sms = sm_count();
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_dstate_tc_kernel, D1_SMEM));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_dx_tc_kernel<BWD_THREADS, false>, D2_SMEM));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_dbc_tc_kernel<BWD_THREADS, false>, D3_SMEM));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_dx_tc_kernel<BWD_THREADS, true>, D2_SMEM));
  HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_dbc_tc_kernel<BWD_THREADS, true>, D3_SMEM));
  int items = ndir * B * H;
  // (three or four CTAs per SM -- 74 KB, 85 registers would fit -- measured no faster: 252.9 / 255.3 / 256.6 us with 3 / 2 / 4)
  hnb::launch_pdl(ssd_bwd_dstate_tc_kernel, dim3(items < 2 * sms ? items : 2 * sms), dim3(D1_THREADS), D1_SMEM, st, tmX, tmDY, p);
  HNB_LAUNCH_CHECK("ssd_bwd_dstate_tc");
  if (variant == 0) {                                   // dx / ddt / dA / dD and dB / dC in one pass
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_fused_tc_kernel<false>, DF_SMEM));
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_bwd_fused_tc_kernel<true>, DF_SMEM));
    items = span_ctas();                                 // one resident CTA per SM: the fix-up wait relies on it
    if (debug) hnb::launch_pdl(ssd_bwd_fused_tc_kernel<true>, dim3(items), dim3(DF_THREADS), DF_SMEM, st, tmX, tmDY, tmS, tmG, p, cuts);
    else hnb::launch_pdl(ssd_bwd_fused_tc_kernel<false>, dim3(items), dim3(DF_THREADS), DF_SMEM, st, tmX, tmDY, tmS, tmG, p, cuts);
    HNB_LAUNCH_CHECK("ssd_bwd_fused_tc");
    if (debug) {
      long long h[512];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, p.dbg, 4096, cudaMemcpyDeviceToHost);
      cudaFree(p.dbg);
      {
        long long mn = 1LL << 60, mx = 0; int imx = 0;
        for (int k = 0; k < items; ++k) { const long long c = h[32 + 2 * k]; if (c < mn) mn = c; if (c > mx) { mx = c; imx = k; } }
        fprintf(stderr, "[ssd_bwd_fused CTA0 pieces] first head-step of a piece %.0f cycles (%lld) | piece end: wait for the accumulation %.0f, dB|dC out %.0f (%lld)\n",
                h[11] ? (double)h[10] / h[11] : 0.0, h[11], h[14] ? (double)h[12] / h[14] : 0.0, h[14] ? (double)h[13] / h[14] : 0.0, h[14]);
        for (int i = 0; i < 7; ++i)
          if (h[400 + 4 * i + 3])
            fprintf(stderr, "[ssd_bwd_fused CTA0 piece %d] dB in registers after %lld cycles, dC after %lld, all stored after %lld (to_fix %lld from_fix %lld)\n",
                    i, h[400 + 4 * i], h[401 + 4 * i], h[402 + 4 * i], h[403 + 4 * i] & 1, (h[403 + 4 * i] >> 1) & 1);
        fprintf(stderr, "[ssd_bwd_fused per-CTA cycles] min %lld max %lld (CTA %d, %lld steps) | CTA0 %lld (%lld steps) CTA%d %lld (%lld steps) CTA%d %lld (%lld steps)\n",
                mn, mx, imx, h[33 + 2 * imx], h[32], h[33], items / 2, h[32 + 2 * (items / 2)], h[33 + 2 * (items / 2)], items - 1,
                h[32 + 2 * (items - 1)], h[33 + 2 * (items - 1)]);
      }
      for (int w = 0; w < 2; ++w) {
        const long long* d = h + 16 * w;
        const double n = d[9] > 0 ? (double)d[9] : 1.0;
        fprintf(stderr, "[ssd_bwd_fused CTA0 %s] full head-steps %lld | cycles: loads + G,R wait %.0f | epi A + TMEM copies %.0f | dot %.0f | "
                "du wait %.0f | epi B loads %.0f | epi B %.0f | step barrier %.0f | sum %.0f\n", w == 0 ? "warp 0 (light rows)" : "warp 15 (heavy rows)",
                d[9], d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[4] / n, d[5] / n, d[6] / n,
                (d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6]) / n);
      }
    }
    return HNB_OK;
  }
  items = ndir * B * nc * (H / p.dxHh);
  if (debug) ssd_bwd_dx_tc_kernel<BWD_THREADS, true><<<items < sms ? items : sms, BWD_THREADS, D2_SMEM, st>>>(tmX, tmDY, tmS, tmG, p);
  else ssd_bwd_dx_tc_kernel<BWD_THREADS, false><<<items < sms ? items : sms, BWD_THREADS, D2_SMEM, st>>>(tmX, tmDY, tmS, tmG, p);
  HNB_LAUNCH_CHECK("ssd_bwd_dx_tc");
  items = ndir * B * nc * dbc_parts;
  if (debug) ssd_bwd_dbc_tc_kernel<BWD_THREADS, true><<<items < sms ? items : sms, BWD_THREADS, D3_SMEM, st>>>(tmX, tmDY, tmS, tmG, p);
  else ssd_bwd_dbc_tc_kernel<BWD_THREADS, false><<<items < sms ? items : sms, BWD_THREADS, D3_SMEM, st>>>(tmX, tmDY, tmS, tmG, p);
  HNB_LAUNCH_CHECK("ssd_bwd_dbc_tc");
  if (debug) {
    long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, p.dbg, 128, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    const double m = h[14] > 0 ? (double)h[14] : 1.0;
    fprintf(stderr, "[ssd_bwd_dbc CTA0] head-steps %lld | cycles/step: wait load %.0f, R mma wait (+Xw,dYs) %.0f, W epilogue %.0f, "
            "mma2 wait %.0f\n", h[14], h[8] / m, h[9] / m, h[10] / m, h[11] / m);
    const double n = h[6] > 0 ? (double)h[6] : 1.0;
    fprintf(stderr, "[ssd_bwd_dx CTA0] items %lld | cycles/item: wait load %.0f, mma1 wait (tables+dot) %.0f, epiA %.0f, "
            "mma2 wait %.0f, epiB+cumsum %.0f\n", h[6], h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n);
  }
  return HNB_OK;
}
