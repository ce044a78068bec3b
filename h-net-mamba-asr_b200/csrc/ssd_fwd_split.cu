// SSD forward on tcgen05, split by dependency instead of by head (impl 5; chosen by impl 1 for rows of >= 8 chunks, where it
// beats the per-(row, head) persistent kernel of ssd_tcgen05.cu -- measured 154 vs 167 us per call on 60 s utterances):
//
//   1. ssd_states_tc_kernel   item = (row, head group), chunks in order.  Only the recurrence that truly is sequential:
//          dS_c = B_c^T (w o X_c)      128x64x128 per head   (B^T: MN-major A operand, as TMA wrote B)
//          S_{c+1} = e^{cs_last} S_c + dS_c               -> states [row, head, chunk, 128 n, 64 p] (bf16, also the
//                                                            backward's operand)
//      The B tile of a chunk is shared by the heads of the item; a chunk step scales the X tiles of its heads in place,
//      issues all their MMAs as one burst and updates the states after one completion wait.
//
//   2. ssd_scan_tc_kernel     item = (row, chunk, head group): every chunk of every row at once.
//          G  = C B^T                   ONCE per item: ngroups = 1, so the score tile is the same for every head;
//                                       it is read from TMEM once and kept in registers (32 fp32 per thread)
//          per head h:  M_h = bf16(G o L_h o dt_h)        registers -> TMEM (A operand, two buffers)
//                       Yd  = M_h X_h,  Yo = C S_in,h     16 MMAs, two accumulator pairs
//                       y   = Yd + e^{cs} Yo + D x        one head behind: runs while the tensor pipe works on the
//                                                          next head
//      One block-wide barrier per head; X | S_in | tables arrive through a three-stage TMA ring two heads ahead; the C
//      tile is double buffered over items and B is refetched as soon as G has retired, so an item's prologue finds
//      its operands in shared memory.
//
// The persistent kernel recomputed G and read its 64 KB from TMEM once per HEAD (12-16 x per chunk) and ran
// load -> G -> epilogue -> Yd/Yo -> epilogue -> dS -> epilogue as one serial chain per chunk step.
#include <cstdlib>

#include "common.cuh"
#include "ssd_tc.cuh"
#include "umma.cuh"

namespace hnb {
namespace {

struct SplitParams {
  const float* Dskip;           // [ndir, H]
  __nv_bfloat16* y;             // [ndir*B*L, di]
  __nv_bfloat16* states;        // [ndir*B, H, nc, 128(n), 64(p)]  state ENTERING each chunk
  const float* tables;          // [ndir*B, H, nc, TAB_FLOATS]
  const __nv_bfloat16* xconv;   // [ndir*B*L, di + 2N]
  int ndirB, B, L, H, di, nc;
  int nh;                       // heads per item
  int n_items;                  // states: ndirB * (H / nh); scan: ndirB * nc * (H / nh)
  int nfc, n_full;              // scan: full chunks per row, ndirB * nfc (full chunks are handed out first)
  FastDiv dHG, dnh, dper, dnfc, dB;
};

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r.x), "r"(r.y), "r"(r.z),
               "r"(r.w)
               : "memory");
}

// ===================================================================================================
// 1. chunk states
// ===================================================================================================
constexpr int ST_THREADS = 512;
constexpr int ST_MAXH = 4;                                     // heads per item (16 fp32 state registers per head and thread)
constexpr int ST_OFF_B = 0;                                    // two buffers (chunk-step parity) of two HALF blocks
constexpr int ST_OFF_X = 4 * HALF;                             // 2 x NH blocks: (chunk-step parity, head)
__host__ __device__ constexpr int st_off_w(int nh) { return ST_OFF_X + 2 * nh * HALF; }       // 2 x NH x (w | ecs) = 1024 B
__host__ __device__ constexpr int st_off_bar(int nh) { return st_off_w(nh) + 2 * nh * 1024; }
__host__ __device__ constexpr int st_smem(int nh) { return st_off_bar(nh) + 128 + 1024; }
static_assert(st_smem(ST_MAXH) <= 232448, "SSD states kernel: shared memory");

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One chunk step = the NH heads of a (row, chunk): their X tiles are scaled in place, ALL their MMAs (8 NH) are issued as one
// burst into NH accumulators, and the state updates read the accumulators after ONE completion wait.  (A tcgen05.ld issued
// while MMAs are in flight completes only after them, and an MMA group costs ~300 cycles of launch + commit latency on top
// of its 8 x 50: one dependent round per HEAD was ~2400 cycles; a round per chunk step amortises that over NH heads, and
// the next chunk step's X tiles are scaled while the burst runs.)
template <int NH>
__global__ void __launch_bounds__(ST_THREADS, 1)
ssd_states_tc_kernel(const __grid_constant__ CUtensorMap tmX, const SplitParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sB = base + ST_OFF_B; uint8_t* sX = base + ST_OFF_X;
  float* sW = reinterpret_cast<float*>(base + st_off_w(NH));
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(base + st_off_bar(NH));   // [2]
  uint64_t* bar_x = bar_b + 2;                                            // [2]: all NH tiles of a chunk step
  uint64_t* bar_m = bar_x + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_m + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2;                 // TMEM lane quarter, 16-column group
  const int row = lq * 32 + lane;                          // state row n
  const bool issuer = tid == 0, loader = tid == 128;
  if (tid == 0) {
    umma::prefetch_tmap(&tmX);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(bar_b + i, 1); umma::mbar_init(bar_x + i, 1); }
    umma::mbar_init(bar_m, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  const int H = p.H, di = p.di, nc = p.nc, ns = nc - 1;
  constexpr uint32_t idesc_s = umma::make_idesc_bf16(128, 64, 1, 1);
  const int my_items = blockIdx.x < p.n_items ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nq = my_items * ns;                            // chunk steps of this CTA

  // ---- loader: cursor over the chunk steps (item, chunk) still to fetch: B | NH x (X, w | ecs)
  int l_k = 0, l_c = 0, l_q = 0;
  auto issue_step = [&]() {
    if (l_k >= my_items) return;
    int db, hg; p.dHG.divmod((int)blockIdx.x + l_k * (int)gridDim.x, db, hg);
    const int par = l_q & 1;
    uint8_t* dst = sB + par * 2 * HALF;
    umma::mbar_expect_tx(bar_b + par, 2 * HALF);
    umma::tma_load_3d(dst, &tmX, bar_b + par, di, l_c * TQ, db);
    umma::tma_load_3d(dst + HALF, &tmX, bar_b + par, di + 64, l_c * TQ, db);
    umma::mbar_expect_tx(bar_x + par, NH * (HALF + 1024));
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const int h = hg * NH + hh, stg = par * NH + hh;
      umma::bulk_load(sW + stg * 256, p.tables + (long long)((db * H + h) * nc + l_c) * TAB_FLOATS + 2 * TQ, 1024, bar_x + par);
      umma::tma_load_3d(sX + stg * HALF, &tmX, bar_x + par, h * TP, l_c * TQ, db);
    }
    ++l_q;
    if (++l_c == ns) { l_c = 0; ++l_k; }
  };
  if (loader) { issue_step(); issue_step(); }

  float S[NH][16];
#pragma unroll
  for (int i = 0; i < NH; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) S[i][j] = 0.f;
  float dec_cur[NH], dec_next[NH];                         // e^{cs_last} per head of the chunk step in flight / the one scaled ahead

  // Xw = w_s X in place for the NH tiles of chunk step qq (the swizzle permutes 16-byte chunks inside a row only)
  auto scale_step = [&](int qq, float* dec) {
    const int par = qq & 1;
    umma::mbar_wait(bar_x + par, (qq >> 1) & 1);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const float* w = sW + (par * NH + hh) * 256;
      dec[hh] = w[TQ + TQ - 1];
      uint4* xs = reinterpret_cast<uint4*>(sX + (par * NH + hh) * HALF);
      const uint4 r0 = xs[tid], r1 = xs[tid + ST_THREADS];
      const float w0 = w[tid >> 3], w1 = w[(tid + ST_THREADS) >> 3];
      float v[8];
      unpack8(r0, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= w0;
      xs[tid] = pack8(v);
      unpack8(r1, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= w1;
      xs[tid + ST_THREADS] = pack8(v);
    }
    umma::fence_async_smem();
  };
  if (nq > 0) scale_step(0, dec_cur);

  int q = 0;
  for (int k = 0; k < my_items; ++k) {
    int db, hg; p.dHG.divmod((int)blockIdx.x + k * (int)gridDim.x, db, hg);
    for (int c = 0; c < ns; ++c, ++q) {
      const int par = q & 1;
      umma::tc_fence_before();
      __syncthreads();                                     // X of this step scaled; the accumulators read by the previous update
      if (issuer) {
        umma::mbar_wait(bar_b + par, (q >> 1) & 1);
        umma::tc_fence_after();
        // descriptors: the start-address field counts 16-byte units, so a step of 2048 B is +128 on the low word
        const uint64_t dB0 = umma::make_smem_desc(umma::smem_u32(sB + par * 2 * HALF), HALF, 1024);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const uint64_t dX0 = umma::make_smem_desc(umma::smem_u32(sX + (par * NH + hh) * HALF), 1024, 1024);
#pragma unroll
          for (int kb = 0; kb < 8; ++kb)                   // dS_h = B^T (w_h o X_h), k = time
            umma::mma_bf16_ss(tmem + 64u * hh, dB0 + (uint64_t)(kb * 128), dX0 + (uint64_t)(kb * 128), idesc_s, kb > 0);
        }
        umma::mma_commit(bar_m);
      }
      if (q + 1 < nq) scale_step(q + 1, dec_next);         // under the burst
      umma::mbar_wait(bar_m, q & 1);
      umma::tc_fence_after();
      if (loader) issue_step();                            // chunk step q + 2 into the buffers the burst has just released
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const int h = hg * NH + hh;
        __nv_bfloat16* sg = p.states + ((((long long)db * H + h) * nc + c) * TN + row) * TP + 16 * cg;
        if (c == 0) {                                      // the state entering chunk 0
          *reinterpret_cast<uint4*>(sg) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sg + 8) = make_uint4(0, 0, 0, 0);
        }
        float ds[16];
        umma::tmem_ld16(t_lane + 64u * hh + 16u * cg, ds);
        umma::tmem_ld_wait();
        const float decay = c == 0 ? 0.f : dec_cur[hh];
#pragma unroll
        for (int j = 0; j < 16; ++j) S[hh][j] = decay * S[hh][j] + ds[j];
        sg += (long long)TN * TP;                          // the state entering chunk c + 1
        *reinterpret_cast<uint4*>(sg) = pack8(S[hh]);
        *reinterpret_cast<uint4*>(sg + 8) = pack8(S[hh] + 8);
      }
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) dec_cur[hh] = dec_next[hh];
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// ===================================================================================================
// 2. chunk scan
// ===================================================================================================
constexpr int SC_THREADS = 512;
constexpr int SC_STAGES = 3;
constexpr int SC_OFF_C = 0;                                    // two buffers (item parity) of two HALF blocks
constexpr int SC_OFF_B = 4 * HALF;                             // two HALF blocks
constexpr int SC_OFF_X = 6 * HALF;                             // SC_STAGES blocks
constexpr int SC_OFF_S = SC_OFF_X + SC_STAGES * HALF;          // SC_STAGES blocks
constexpr int SC_OFF_TAB = SC_OFF_S + SC_STAGES * HALF;
constexpr int SC_OFF_BAR = SC_OFF_TAB + SC_STAGES * TAB_BYTES;
constexpr int SC_SMEM = SC_OFF_BAR + 128 + 1024;
static_assert(SC_SMEM <= 232448, "SSD scan kernel: shared memory");
static_assert(TAB_BYTES % 16 == 0 && SC_OFF_BAR % 8 == 0, "SSD scan kernel: alignment");

__global__ void __launch_bounds__(SC_THREADS, 1)
ssd_scan_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS, const SplitParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* sC = base + SC_OFF_C; uint8_t* sB = base + SC_OFF_B; uint8_t* sX = base + SC_OFF_X; uint8_t* sS = base + SC_OFF_S;
  float* tabs = reinterpret_cast<float*>(base + SC_OFF_TAB);
  uint64_t* bar_c = reinterpret_cast<uint64_t*>(base + SC_OFF_BAR);   // [2]
  uint64_t* bar_b = bar_c + 2;
  uint64_t* bar_g = bar_b + 1;
  uint64_t* bar_h = bar_g + 1;                                        // [SC_STAGES]
  uint64_t* bar_y = bar_h + SC_STAGES;                                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_y + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2;                 // TMEM lane quarter = row block I; column group
  const int row = lq * 32 + lane;
  const bool issuer = tid == 0, loader = tid == 128;       // warps 0 and 4 own the lightest rows of the score tile
  if (tid == 0) {
    umma::prefetch_tmap(&tmX); umma::prefetch_tmap(&tmS);
    umma::mbar_init(bar_c, 1); umma::mbar_init(bar_c + 1, 1); umma::mbar_init(bar_b, 1); umma::mbar_init(bar_g, 1);
    for (int i = 0; i < SC_STAGES; ++i) umma::mbar_init(bar_h + i, 1);
    umma::mbar_init(bar_y, 1); umma::mbar_init(bar_y + 1, 1);
    umma::fence_barrier_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(lq * 32) << 16);
  constexpr uint32_t TM_G = 0, TM_M = 128, TM_YD = 256, TM_YO = 384;          // M, Yd, Yo: two buffers of 64 columns
  const int H = p.H, L = p.L, di = p.di, nc = p.nc, nh = p.nh, C = di + 2 * TN;
  constexpr uint32_t idesc_g = umma::make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_y = umma::make_idesc_bf16(128, 64, 0, 1);
  const int my_items = blockIdx.x < p.n_items ? (p.n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // the blocks of M above the diagonal are zero for every head of every item: written once
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k > lq) { tmem_st4(t_lane + TM_M + 16u * k + 4u * cg, z); tmem_st4(t_lane + TM_M + 64u + 16u * k + 4u * cg, z); }
    umma::tmem_st_wait();
  }

  auto item_of = [&](int k, int& db, int& c, int& hg) {    // k-th item of this CTA: full chunks first
    int base_; p.dHG.divmod((int)blockIdx.x + k * (int)gridDim.x, base_, hg);
    if (base_ < p.n_full) p.dnfc.divmod(base_, db, c); else { db = base_ - p.n_full; c = nc - 1; }
  };
  auto load_cb = [&](int k) {                              // C into buffer k & 1, B into the single B buffer
    int db, c, hg; item_of(k, db, c, hg);
    uint8_t* dst = sC + (k & 1) * 2 * HALF;
    umma::mbar_expect_tx(bar_c + (k & 1), 2 * HALF);
    umma::tma_load_3d(dst, &tmX, bar_c + (k & 1), di + TN, c * TQ, db);
    umma::tma_load_3d(dst + HALF, &tmX, bar_c + (k & 1), di + TN + 64, c * TQ, db);
    umma::mbar_expect_tx(bar_b, 2 * HALF);
    umma::tma_load_3d(sB, &tmX, bar_b, di, c * TQ, db);
    umma::tma_load_3d(sB + HALF, &tmX, bar_b, di + 64, c * TQ, db);
  };
  // loader: cursor over the head steps (item, head) still to fetch: X | S_in | tables
  int lh_k = 0, lh_h = 0, lh_st = 0, lh_db = 0, lh_c = 0, lh_hg = 0;
  if (loader && my_items > 0) item_of(0, lh_db, lh_c, lh_hg);
  auto issue_h = [&]() {
    if (lh_k >= my_items) return;
    const int h = lh_hg * nh + lh_h;
    const int sidx = (lh_db * H + h) * nc + lh_c;
    umma::mbar_expect_tx(bar_h + lh_st, (lh_c > 0 ? 2 * HALF : HALF) + TAB_BYTES);
    umma::bulk_load(tabs + lh_st * TAB_FLOATS, p.tables + (long long)sidx * TAB_FLOATS, TAB_BYTES, bar_h + lh_st);
    umma::tma_load_3d(sX + lh_st * HALF, &tmX, bar_h + lh_st, h * TP, lh_c * TQ, lh_db);
    if (lh_c > 0) umma::tma_load_2d(sS + lh_st * HALF, &tmS, bar_h + lh_st, 0, sidx * TN);
    lh_st = lh_st == SC_STAGES - 1 ? 0 : lh_st + 1;
    if (++lh_h == nh) { lh_h = 0; if (++lh_k < my_items) item_of(lh_k, lh_db, lh_c, lh_hg); }
  };
  if (loader && my_items > 0) { load_cb(0); issue_h(); issue_h(); issue_h(); }

  float G[32];                                             // G[8 k + j] = (C B^T)[row, 32 k + 8 cg + j]
  // what the epilogue of the PREVIOUS head step needs (it runs one step behind)
  float ecs_prev = 0.f, Dh_prev = 0.f;
  __nv_bfloat16* y_prev = nullptr;
  uint4 x_prev0 = make_uint4(0, 0, 0, 0), x_prev1 = x_prev0;
  int qv_prev = 0;
  bool yo_prev = false;
  int g = 0, st = 0;                                       // head steps done by this CTA / their stage

  // y of head step g - 1 = Yd + e^{cs} Yo + D x, in two halves around the barrier that releases the next MMAs: tcgen05.ld and
  // tcgen05.mma share one in-order queue, so the TMEM loads are issued BEFORE the next head's MMAs and consumed while they run.
  float yd[16], yo[16];
  auto epi_issue = [&]() {
    const int gp = g - 1, pb = gp & 1;
    umma::mbar_wait(bar_y + pb, (gp >> 1) & 1);
    umma::tc_fence_after();
    if ((row >> 5) < ((qv_prev + 31) >> 5)) {
      umma::tmem_ld16(t_lane + TM_YD + 64u * pb + 16u * cg, yd);
      if (yo_prev) umma::tmem_ld16(t_lane + TM_YO + 64u * pb + 16u * cg, yo);
    }
  };
  auto epi_finish = [&]() {
    umma::tmem_ld_wait();
    if (row < qv_prev) {
      float x[16], o[16];
      unpack8(x_prev0, x); unpack8(x_prev1, x + 8);
      if (yo_prev) {
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = yd[e] + ecs_prev * yo[e] + Dh_prev * x[e];
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = yd[e] + Dh_prev * x[e];
      }
      *reinterpret_cast<uint4*>(y_prev) = pack8(o);
      *reinterpret_cast<uint4*>(y_prev + 8) = pack8(o + 8);
    }
  };

  for (int k = 0; k < my_items; ++k) {
    int db, c, hg; item_of(k, db, c, hg);
    const int q0 = c * TQ, qv = min(TQ, L - q0);
    const int nblk = (qv + 31) >> 5, nkb = (qv + 15) >> 4;
    const uint64_t dC0 = umma::make_smem_desc(umma::smem_u32(sC + (k & 1) * 2 * HALF), 16, 1024);
    const float* Dk = p.Dskip + p.dB.div(db) * H + hg * nh;
    __nv_bfloat16* y_item = p.y + ((long long)db * L + q0 + row) * di + hg * nh * TP + 16 * cg;
    // ---- item prologue: G = C B^T queued behind the previous item's last TMEM loads; that item's last epilogue runs meanwhile
    if (g > 0) {
      epi_issue();
      umma::tc_fence_before();
      __syncthreads();
      // stage of head step g - 1: X, S consumed by its MMAs (epi_issue waited for them), its tables read before that step's
      // barrier, and -- this barrier -- every thread has taken its x for the D x term out of it
      if (loader) issue_h();
    }
    if (issuer) {
      umma::mbar_wait(bar_c + (k & 1), (k >> 1) & 1);
      umma::mbar_wait(bar_b, k & 1);
      umma::tc_fence_after();
      const uint64_t dB0 = umma::make_smem_desc(umma::smem_u32(sB), 16, 1024);
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        const uint64_t o = (uint64_t)(((kb >> 2) * HALF + (kb & 3) * 32) >> 4);
        umma::mma_bf16_ss(tmem + TM_G, dC0 + o, dB0 + o, idesc_g, kb > 0);
      }
      umma::mma_commit(bar_g);
    }
    if (g > 0) epi_finish();
    umma::mbar_wait(bar_g, k & 1);
    umma::tc_fence_after();
    if (loader && k + 1 < my_items) load_cb(k + 1);        // B is dead; the other C buffer retired an item ago
    if (lq < nblk) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        if (kk <= lq) umma::tmem_ld8(t_lane + TM_G + 32u * kk + 8u * cg, G + 8 * kk);
      umma::tmem_ld_wait();
    }
    for (int hh = 0; hh < nh; ++hh) {
      const int mb = g & 1;
      const float Dh = __ldg(Dk + hh);
      // ---- M_h[t, s] = G[t, s] e^{cs_t - cs_s} dt_s  (s <= t)  -> bf16 pairs in TMEM
      umma::mbar_wait(bar_h + st, (g / SC_STAGES) & 1);
      const float* tab = tabs + st * TAB_FLOATS;
      const float ecs_cur = tab[3 * TQ + row];
      if (lq < nblk) {
        const int t = row, I = lq;
        const float cs_t = tab[t], e_ref = I > 0 ? __expf(cs_t - tab[32 * I - 1]) : 0.f;
        const float cs2 = cs_t * 1.4426950408889634f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (kk <= I) {
            const int s0 = 32 * kk + 8 * cg;
            float l[8], d8[8], m8[8];
            load8(d8, tab + TQ + s0);
            if (kk < I) {
              load8(l, tab + 5 * TQ + I * TQ + s0);
#pragma unroll
              for (int j = 0; j < 8; ++j) m8[j] = G[8 * kk + j] * (e_ref * (l[j] * d8[j]));
            } else {
              load8(l, tab + s0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float e = ex2_approx(fmaf(l[j], -1.4426950408889634f, cs2));
                m8[j] = (s0 + j <= t) ? G[8 * kk + j] * (e * d8[j]) : 0.f;
              }
            }
            tmem_st4(t_lane + TM_M + 64u * mb + 16u * kk + 4u * cg, pack8(m8));
          }
        umma::tmem_st_wait();
      }
      if (hh > 0) epi_issue();
      umma::tc_fence_before();
      __syncthreads();
      if (loader && hh > 0) issue_h();                     // refill the stage of head step g - 1 (see the item prologue)
      if (issuer) {
        umma::tc_fence_after();
        const uint64_t dX0 = umma::make_smem_desc(umma::smem_u32(sX + st * HALF), 1024, 1024);
        const uint64_t dS0 = umma::make_smem_desc(umma::smem_u32(sS + st * HALF), 1024, 1024);
        const uint32_t tYd = tmem + TM_YD + 64u * mb, tYo = tmem + TM_YO + 64u * mb, tM = tmem + TM_M + 64u * mb;
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)                     // Yd = M X   (k = time: valid frames only)
          if (kb < nkb) umma::mma_bf16_ts(tYd, tM + 8u * kb, dX0 + (uint64_t)(kb * 128), idesc_y, kb > 0);
        if (c > 0) {
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {                 // Yo = C S_in
            const uint64_t o = (uint64_t)(((kb >> 2) * HALF + (kb & 3) * 32) >> 4);
            umma::mma_bf16_ss(tYo, dC0 + o, dS0 + (uint64_t)(kb * 128), idesc_y, kb > 0);
          }
        }
        umma::mma_commit(bar_y + mb);
      }
      if (hh > 0) epi_finish();
      // hand this head step to the epilogue that runs one step behind; x for the D x term comes from the X tile that the
      // MMAs just queued are reading (the stage is refilled only after they retire)
      {
        const uint8_t* xs = sX + st * HALF;
        x_prev0 = *reinterpret_cast<const uint4*>(xs + swz(row, 2 * cg));
        x_prev1 = *reinterpret_cast<const uint4*>(xs + swz(row, 2 * cg + 1));
      }
      ecs_prev = ecs_cur; Dh_prev = Dh;
      y_prev = y_item + hh * TP;
      qv_prev = qv; yo_prev = c > 0;
      ++g;
      st = st == SC_STAGES - 1 ? 0 : st + 1;
    }
  }
  if (g > 0) { epi_issue(); epi_finish(); }                // (nothing left to fetch)
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int sm_count_split() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Heads per item: more heads amortise the item prologue (C | B loads, G, 64 KB of TMEM reads: ~2.5 head steps), fewer
// heads balance the static round-robin better.  Cost unit: one head step of a full chunk.
int scan_heads_per_item(int ndirB, int L, int H, int sms) {
  const int nc = cdiv(L, TQ), rem = L - (nc - 1) * TQ;
  const long long nfull = (long long)ndirB * (rem == TQ ? nc : nc - 1), npart = (long long)ndirB * nc - nfull;
  int best = 1;
  double best_cost = 1e30;
  for (int nh = 1; nh <= H; ++nh) {
    if (H % nh) continue;
    const int groups = H / nh;
    const double full = nh + 2.5, part = (0.35 + 0.65 * rem / TQ) * nh + 2.5;
    double cost = 0.0;
    for (long long i = 0; i < (nfull + npart) * groups; i += sms) cost += i < nfull * groups ? full : part;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = nh; }
  }
  return best;
}

}  // namespace
}  // namespace hnb

using namespace hnb;

// tables (built by the caller) -> states -> y.  `states` is the tcgen05 workspace: states followed by the tables.
int hnb_ssd_fwd_split_tc(const CUtensorMap* tmX, const void* xconv, const float* Dskip, const float* tables, int ndir, int B,
                         int L, int di, int H, void* y, void* states, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = cdiv(L, TQ), sms = sm_count_split();
  SplitParams p;
  p.Dskip = Dskip; p.y = (__nv_bfloat16*)y; p.states = (__nv_bfloat16*)states; p.tables = tables;
  p.xconv = (const __nv_bfloat16*)xconv;
  p.ndirB = ndir * B; p.B = B; p.L = L; p.H = H; p.di = di; p.nc = nc;
  p.dB = FastDiv(B);
  p.nfc = (L % TQ == 0) ? nc : nc - 1; p.n_full = p.ndirB * p.nfc;
  p.dnfc = FastDiv(p.nfc > 0 ? p.nfc : 1);
  CUtensorMap tmS;
  {
    uint64_t d2[2] = {(uint64_t)TP, (uint64_t)ndir * B * H * nc * TN};
    uint64_t s2[1] = {(uint64_t)TP * 2};
    uint32_t b2[2] = {TP, TN};
    int rc = make_tmap_bf16(&tmS, states, 2, d2, s2, b2);
    if (rc) return rc;
  }
  if (nc == 1) {
    HNB_CUDA_CALL(cudaMemsetAsync(states, 0, (size_t)ndir * B * H * TN * TP * sizeof(__nv_bfloat16), st));
  } else {
    static const int force = getenv("HNB_SSD_STATE_HEADS") ? atoi(getenv("HNB_SSD_STATE_HEADS")) : 0;
    int nh = ST_MAXH;
    while (H % nh) --nh;
    if (force > 0 && force <= ST_MAXH && H % force == 0) nh = force;
    p.nh = nh; p.n_items = p.ndirB * (H / nh);
    p.dHG = FastDiv(H / nh); p.dnh = FastDiv(nh); p.dper = FastDiv((nc - 1) * nh);
    const int grid = p.n_items < sms ? p.n_items : sms;
#define HNB_ST_LAUNCH(NH_)                                                                               \
    do {                                                                                                 \
      HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_states_tc_kernel<NH_>, st_smem(NH_)));            \
      hnb::launch_pdl(ssd_states_tc_kernel<NH_>, dim3(grid), dim3(ST_THREADS), st_smem(NH_), st, *tmX, p);                         \
    } while (0)
    if (nh == 4) HNB_ST_LAUNCH(4); else if (nh == 3) HNB_ST_LAUNCH(3); else if (nh == 2) HNB_ST_LAUNCH(2); else HNB_ST_LAUNCH(1);
#undef HNB_ST_LAUNCH
    HNB_LAUNCH_CHECK("ssd_states_tc");
  }
  {
    static const int force = getenv("HNB_SSD_SCAN_HEADS") ? atoi(getenv("HNB_SSD_SCAN_HEADS")) : 0;
    const int nh = (force > 0 && H % force == 0) ? force : scan_heads_per_item(p.ndirB, L, H, sms);
    p.nh = nh; p.n_items = p.ndirB * nc * (H / nh);
    p.dHG = FastDiv(H / nh); p.dnh = FastDiv(nh); p.dper = FastDiv(1);
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)ssd_scan_tc_kernel, SC_SMEM));
    hnb::launch_pdl(ssd_scan_tc_kernel, dim3(p.n_items < sms ? p.n_items : sms), dim3(SC_THREADS), SC_SMEM, st, *tmX, tmS, p);
    HNB_LAUNCH_CHECK("ssd_scan_tc");
  }
  return HNB_OK;
}
