// bf16 GEMM on tcgen05 tensor cores: TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma (128 x 256 x 16,
// fp32 accumulators in TMEM) -> tcgen05.ld epilogue with fused bias / residual / dtype cast / split-K atomics.
// Serves the in/out projections of the mixer (both directions fused into one GEMM each), the router
// W_q|W_k projection, proj_in/proj_out, and every dgrad / wgrad of those.
//
//   C[M,N] = A(M,K) . B(N,K)^T      A: K-major [M,K] or MN-major [K,M];  B: K-major [N,K] or MN-major [K,N]
//
// Persistent and warp-specialised: one CTA per SM walks 128 x 256 (or 128 x 128) output tiles; warp 0 drives TMA
// through a 4-6 stage smem ring, warp 1 issues the MMAs into one of two TMEM accumulators, eight epilogue warps
// drain the other accumulator (TMEM lane quarter = warp % 4), so the epilogue of tile i overlaps the main loop of
// tile i+1.  The projections of this model have K = 384..512 (6-8 k-blocks per tile): without that overlap the
// prologue/epilogue, not the tensor pipe, sets the pace.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace hnb {

PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable"); return HNB_ERR_CUDA; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base %p dims %llu,%llu stride %llu box %u,%u", (int)r, base,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    return HNB_ERR_CUDA;
  }
  return HNB_OK;
}

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_STAGE = BM * BK * 2;                          // 16 KB
constexpr int GEMM_THREADS = 64 + 256;                        // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int EPI_WARPS = 8;

template <int BN> struct GemmCfg {
  static constexpr int B_STAGE = BN * BK * 2;                 // 16 / 32 KB
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 7);   // as many bytes in flight as 227 KB allows
  static constexpr int TMEM_COLS = BN == 192 ? 512 : 2 * BN;           // two accumulators; a power of two
  static constexpr int RING = STAGES * (A_STAGE + B_STAGE);
  static constexpr int SMEM = RING + 256 + 1024;
};

// 256-bit global accesses (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// descriptor = (hi << 32) | lo with lo = start address (16-byte units, 14 bits) | LBO << 16 and a constant hi word (SBO 1024 B,
// version 1, SWIZZLE_128B): the issuing thread only adds to the 32-bit low word (uniform datapath), see the MMA loops
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(0x40004040u));
  return d;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}

struct GemmParams {
  int M, N, K;
  int tiles_m, tiles_n, splitk, kblocks_per_split, total_kb;
  int groups_m;               // tiles_m / cluster size (rounded up): one work item = CL vertically adjacent tiles
  const float* bias;
  const void* R;
  long long ldr;
  void* C;
  long long ldc;
  int c_is_f32;
  int atomic;
  int cblk;                   // > 0: C is stored as N / cblk separate [M, cblk] matrices (fp32 outputs; cblk % 32 == 0)
  long long cblk_stride;      // elements between those matrices
  // element offset of (row, col0) in C; a 32-column block never straddles two column blocks
  __device__ __forceinline__ long long c_off(int row, int col0) const {
    if (cblk <= 0) return (long long)row * ldc + col0;
    const int blk = col0 / cblk;
    return blk * cblk_stride + (long long)row * ldc + (col0 - blk * cblk);
  }
};

// Epilogue of one accumulator tile for one warp: TMEM lane quarter at `t_addr`, output row `row`, the warp's column
// half `chalf` of the BN-wide tile starting at column n0.
// `release` is the accumulator's tmem_empty barrier: the warp arrives on it (on the pair leader's copy when PAIR) as soon
// as its last TMEM load has landed in registers, i.e. before the last block is converted and stored.
template <int BN, bool PAIR>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t t_addr, int row, int n0, int chalf,
                                              uint64_t* release, int lane) {
      // TMEM hands each lane one ROW: 32 consecutive columns = 64 B (bf16) / 128 B (fp32) of that row.  They leave as
      // 256-bit stores, one full 32-byte sector per lane and instruction -- no shared-memory transpose (the first
      // version spent 12 k warp-instructions per tile in one: the K = 384 projections were epilogue-bound).  The TMEM
      // load of block j+1 is in flight while block j is converted and stored (two register buffers).
      constexpr int NBLK = BN / 64;
      const int cbase = chalf * (BN / 2);
      float vbuf[2][32];
      bool released = false;
      if (n0 + cbase < p.N) umma::tmem_ld32(t_addr + (uint32_t)cbase, vbuf[0]);
#pragma unroll
      for (int jb = 0; jb < NBLK; ++jb) {
        const int c0 = cbase + 32 * jb;
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;                             // warp-uniform
        umma::tmem_ld_wait();
        const bool more = jb + 1 < NBLK && col0 + 32 < p.N;
        if (more) {
          umma::tmem_ld32(t_addr + (uint32_t)(c0 + 32), vbuf[(jb + 1) & 1]);
        } else {
          umma::tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) umma::mbar_arrive_leader(release); else umma::mbar_arrive(release); }
          released = true;
        }
        float* v = vbuf[jb & 1];
        if (row < p.M) {
        if (p.bias) {
          if (col0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + col0 + j);       // warp-uniform addresses: broadcast
          } else {
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        if (p.atomic) {
          float* c = reinterpret_cast<float*>(p.C) + p.c_off(row, col0);
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(c) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
          } else {
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(c + j, v[j]);
          }
        } else if (p.c_is_f32) {
          float* c = reinterpret_cast<float*>(p.C) + p.c_off(row, col0);
          const float* r = p.R ? reinterpret_cast<const float*>(p.R) + (long long)row * p.ldr + col0 : nullptr;
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(c) & 31) == 0 && (!r || (reinterpret_cast<uintptr_t>(r) & 31) == 0)) {
            if (r) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float t[8];
                ldg256(r + j, reinterpret_cast<uint32_t*>(t));
#pragma unroll
                for (int e = 0; e < 8; ++e) v[j + e] += t[e];
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) stg256(c + j, reinterpret_cast<const uint32_t*>(v + j));
          } else {
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) c[j] = v[j] + (r ? r[j] : 0.f);
          }
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + col0;
          const __nv_bfloat16* r = p.R ? reinterpret_cast<const __nv_bfloat16*>(p.R) + (long long)row * p.ldr + col0 : nullptr;
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(c) & 31) == 0 && (!r || (reinterpret_cast<uintptr_t>(r) & 31) == 0)) {
            if (r) {
#pragma unroll
              for (int j = 0; j < 32; j += 16) {
                uint32_t t[8];
                ldg256(r + j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t[e]));
                  v[j + 2 * e] += f.x; v[j + 2 * e + 1] += f.y;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 16) {
              uint32_t o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
                o[e] = *reinterpret_cast<const uint32_t*>(&h);
              }
              stg256(c + j, o);
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) c[j] = __float2bfloat16_rn(v[j] + (r ? __bfloat162float(r[j]) : 0.f));
          }
        }
        }
        __syncwarp();                                       // tcgen05.ld is warp-collective: reconverge before the next one
      }
      if (!released) {                                      // tile entirely beyond N for this warp's column half
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) umma::mbar_arrive_leader(release); else umma::mbar_arrive(release); }
      }
}

// Persistent, warp-specialised: each CTA walks work items (k-split, n-tile, m-tile) with a static stride.
//   TMA warp  -> smem ring (full/empty mbarriers)
//   MMA warp  -> tcgen05.mma 128 x BN x 16 into one of TWO TMEM accumulators (tmem_full/tmem_empty mbarriers)
//   8 epilogue warps drain accumulator i while the MMA warp already fills accumulator i^1.
//
// CL = 2: the two CTAs of a cluster take vertically adjacent tiles of the same column block, so they need the same B
// tile.  Each loads HALF of it and multicasts that half into both CTAs' shared memory: 32 KB instead of 48 KB of L2
// reads per CTA and k-block (these K = 384..512 GEMMs are paced by L2 -> SM bandwidth, not by the tensor pipe).  The
// MMAs stay 1-SM; a stage is refilled only when BOTH CTAs' MMAs have released it (multicast commit, empty count 2).
template <int TA, int TB, int BN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  pdl_trigger();
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, B_STAGE = Cfg::B_STAGE;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::RING);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;                       // [2]
  uint64_t* tmem_empty = tmem_full + 2;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.groups_m * p.tiles_n * p.splitk;
  const int crank = CL > 1 ? (int)umma::cluster_ctarank() : 0;
  const int first_item = blockIdx.x / CL, item_stride = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA);
    umma::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], CL); }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&tmem_full[a], 1); umma::mbar_init(&tmem_empty[a], EPI_WARPS); }
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  umma::tc_fence_before();
  __syncthreads();
  if (CL > 1) umma::cluster_sync_all();        // the peer's barriers are initialised before anything is multicast into them
  umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& nkb) {
    // column tile fastest: the CTAs working at the same time share their A rows (read from DRAM once, then L2 hits);
    // B is a weight matrix or a narrow activation and stays in L2 anyway.  (Row-fastest order re-read the 115 MB
    // dzxbcdt operand of the in-projection dgrad once per column tile: ncu, 207 MB of DRAM reads.)
    const int tn = item % p.tiles_n;
    const int r = item / p.tiles_n;
    const int tm = r % p.groups_m, ks = r / p.groups_m;
    m0 = (tm * CL + crank) * BM; n0 = tn * BN;          // may start beyond M for the odd last tile: loads zero-fill
    kb0 = ks * p.kblocks_per_split;
    nkb = min(p.kblocks_per_split, p.total_kb - kb0);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t it = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        int m0, n0, kb0, nkb;
        decode(item, m0, n0, kb0, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          umma::mbar_wait(&empty[s], ph ^ 1);
          umma::mbar_expect_tx(&full[s], A_STAGE + B_STAGE);
          const int k0 = (kb0 + i) * BK;
          uint8_t* a = sA + s * A_STAGE;
          uint8_t* b = sB + s * B_STAGE;
          if (TA == 0) {
            umma::tma_load_2d(a, &tmA, &full[s], k0, m0);                  // box {64 k, 128 m}
          } else {
            umma::tma_load_2d(a, &tmA, &full[s], m0, k0);                  // boxes {64 m, 64 k}
            umma::tma_load_2d(a + 8192, &tmA, &full[s], m0 + 64, k0);
          }
          if (CL == 1) {
            if (TB == 0) {
              umma::tma_load_2d(b, &tmB, &full[s], k0, n0);                // box {64 k, BN n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) umma::tma_load_2d(b + j * 8192, &tmB, &full[s], n0 + 64 * j, k0);
            }
          } else {                                                         // this CTA's half of B, into both CTAs
            if (TB == 0) {
              umma::tma_load_2d_mc(b + crank * (B_STAGE / 2), &tmB, &full[s], k0, n0 + crank * (BN / 2), 0x3);   // box {64 k, BN/2 n}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = crank * (BN / 128) + j;
                umma::tma_load_2d_mc(b + jj * 8192, &tmB, &full[s], n0 + 64 * jj, k0, 0x3);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (umma::elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma::make_idesc_bf16(BM, BN, TA, TB);
      const uint32_t dA_ring = desc_lo(umma::smem_u32(sA), TA == 0 ? 16 : 8192);
      const uint32_t dB_ring = desc_lo(umma::smem_u32(sB), TB == 0 ? 16 : 8192);
      uint32_t it = 0, li = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++li) {
        int m0, n0, kb0, nkb;
        decode(item, m0, n0, kb0, nkb);
        const uint32_t acc = li & 1, aph = (li >> 1) & 1;
        umma::mbar_wait(&tmem_empty[acc], aph ^ 1);                        // the epilogue has drained this accumulator
        umma::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          umma::mbar_wait(&full[s], ph);
          umma::tc_fence_after();
          // descriptors by 64-bit adds on the ring's base descriptors (the start-address field counts 16-byte units and
          // never carries out of its 14 bits: shared memory ends below 256 KB).  The issuing thread is ONE thread: rebuilding
          // both descriptors from addresses cost ~16 instructions per MMA and made the issue loop (~830 cycles per k-block,
          // HNB_GEMM_DEBUG=1) slower than the 512 cycles the four MMAs take.
          const uint32_t da0 = dA_ring + (uint32_t)s * (uint32_t)(A_STAGE >> 4);
          const uint32_t db0 = dB_ring + (uint32_t)s * (uint32_t)(B_STAGE >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma::mma_bf16_ss(tmem_d, desc_from_lo(da0 + (uint32_t)(TA == 0 ? k * 2 : k * 128)),
                              desc_from_lo(db0 + (uint32_t)(TB == 0 ? k * 2 : k * 128)), idesc, (i > 0 || k > 0) ? 1u : 0u);
          if (CL == 1) umma::mma_commit(&empty[s]);                        // frees the smem slot when the MMAs retire
          else umma::mma_commit_mc(&empty[s], 0x3);                        // ... in both CTAs: either may refill it
        }
        umma::mma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    const int ew = warp - 2;                      // 0..7
    const int lq = warp & 3;                      // TMEM lane quarter this warp may read
    const int chalf = ew >> 2;                    // column half of the tile
    uint32_t li = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++li) {
      int m0, n0, kb0, nkb;
      decode(item, m0, n0, kb0, nkb);
      const uint32_t acc = li & 1, aph = (li >> 1) & 1;
      umma::mbar_wait(&tmem_full[acc], aph);
      umma::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16);
      epilogue_tile<BN, false>(p, t_addr, m0 + lq * 32 + lane, n0, chalf, &tmem_empty[acc], lane);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (CL > 1) umma::cluster_sync_all();        // no CTA leaves while its peer can still signal or fill its shared memory
  if (warp == 1) umma::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---- CTA-pair variant (cta_group::2): 256 x 256 tiles over two SMs --------------------------------------------
// The long-K GEMMs are paced by TMA latency x bytes in flight per SM.  With the pair MMA each CTA stages only ITS 128
// rows of A and ITS 128 of the tile's 256 B rows per k-block (32 KB instead of 48 KB), so six stages fit and the
// same latency covers 1.5x the k-blocks.  Rank 0 ("leader") issues every MMA; both CTAs run a TMA producer whose bytes
// are counted on the LEADER's full barrier, the MMA's completion is multicast to both CTAs' empty / tmem_full
// barriers, both CTAs drain their own 128 accumulator rows, and the peer's epilogue warps arrive on the leader's
// tmem_empty barrier.
constexpr int PAIR_STAGES = 6, PAIR_BN = 256, PAIR_B_STAGE = 128 * BK * 2;
constexpr int PAIR_RING = PAIR_STAGES * (A_STAGE + PAIR_B_STAGE);
constexpr int PAIR_SMEM = PAIR_RING + 256 + 1024;

template <int TA, int TB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + PAIR_STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PAIR_RING);
  uint64_t* empty = full + PAIR_STAGES;
  uint64_t* tmem_full = empty + PAIR_STAGES;                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                       // [2]  (the leader's copy is the one that counts)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.groups_m * p.tiles_n * p.splitk;
  const int crank = (int)umma::cluster_ctarank();
  const int first_item = blockIdx.x / 2, item_stride = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA);
    umma::prefetch_tmap(&tmB);
    for (int s = 0; s < PAIR_STAGES; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&tmem_full[a], 1); umma::mbar_init(&tmem_empty[a], 2 * EPI_WARPS); }
    umma::fence_barrier_init();
  }
  if (warp == 1) umma::tmem_alloc_pair(tmem_slot, 2 * PAIR_BN);
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  umma::tc_fence_after();
  pdl_wait();                                             // everything above overlaps the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int item, int& m0, int& n0, int& kb0, int& nkb) {
    const int tn = item % p.tiles_n;                           // column tile fastest, as above
    const int r = item / p.tiles_n;
    const int tm = r % p.groups_m, ks = r / p.groups_m;
    m0 = tm * 2 * BM; n0 = tn * PAIR_BN;                       // the PAIR's 256 x 256 tile
    kb0 = ks * p.kblocks_per_split;
    nkb = min(p.kblocks_per_split, p.total_kb - kb0);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): my 128 rows of A, my 128 of the tile's B rows ----------------
      uint32_t it = 0;
      for (int item = first_item; item < n_items; item += item_stride) {
        int m0, n0, kb0, nkb;
        decode(item, m0, n0, kb0, nkb);
        const int ma = m0 + crank * BM, nb = n0 + crank * 128;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % PAIR_STAGES;
          const uint32_t ph = (it / PAIR_STAGES) & 1;
          umma::mbar_wait(&empty[s], ph ^ 1);
          if (crank == 0) umma::mbar_expect_tx(&full[s], 2 * (A_STAGE + PAIR_B_STAGE));     // both CTAs' bytes
          const int k0 = (kb0 + i) * BK;
          uint8_t* a = sA + s * A_STAGE;
          uint8_t* b = sB + s * PAIR_B_STAGE;
          if (TA == 0) {
            umma::tma_load_2d_pair(a, &tmA, &full[s], k0, ma);
          } else {
            umma::tma_load_2d_pair(a, &tmA, &full[s], ma, k0);
            umma::tma_load_2d_pair(a + 8192, &tmA, &full[s], ma + 64, k0);
          }
          if (TB == 0) {
            umma::tma_load_2d_pair(b, &tmB, &full[s], k0, nb);
          } else {
            umma::tma_load_2d_pair(b, &tmB, &full[s], nb, k0);
            umma::tma_load_2d_pair(b + 8192, &tmB, &full[s], nb + 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && umma::elect_one()) {
      // ---------------- MMA issuer (leader only) ----------------
      constexpr uint32_t idesc = umma::make_idesc_bf16(2 * BM, PAIR_BN, TA, TB);
      const uint32_t dA_ring = desc_lo(umma::smem_u32(sA), TA == 0 ? 16 : 8192);
      const uint32_t dB_ring = desc_lo(umma::smem_u32(sB), TB == 0 ? 16 : 8192);
      uint32_t it = 0, li = 0;
      for (int item = first_item; item < n_items; item += item_stride, ++li) {
        int m0, n0, kb0, nkb;
        decode(item, m0, n0, kb0, nkb);
        const uint32_t acc = li & 1, aph = (li >> 1) & 1;
        umma::mbar_wait(&tmem_empty[acc], aph ^ 1);                        // both CTAs have drained this accumulator
        umma::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PAIR_BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % PAIR_STAGES;
          const uint32_t ph = (it / PAIR_STAGES) & 1;
          umma::mbar_wait(&full[s], ph);
          umma::tc_fence_after();
          const uint32_t da0 = dA_ring + (uint32_t)s * (uint32_t)(A_STAGE >> 4);
          const uint32_t db0 = dB_ring + (uint32_t)s * (uint32_t)(PAIR_B_STAGE >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma::mma_bf16_ss_pair(tmem_d, desc_from_lo(da0 + (uint32_t)(TA == 0 ? k * 2 : k * 128)),
                                   desc_from_lo(db0 + (uint32_t)(TB == 0 ? k * 2 : k * 128)), idesc, (i > 0 || k > 0) ? 1u : 0u);
          umma::mma_commit_pair(&empty[s], 0x3);                           // frees the slot in both CTAs
        }
        umma::mma_commit_pair(&tmem_full[acc], 0x3);
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): my 128 rows of the accumulator ----------------
    const int lq = warp & 3, chalf = (warp - 2) >> 2;
    uint32_t li = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++li) {
      int m0, n0, kb0, nkb;
      decode(item, m0, n0, kb0, nkb);
      const uint32_t acc = li & 1, aph = (li >> 1) & 1;
      umma::mbar_wait(&tmem_full[acc], aph);
      umma::tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * PAIR_BN + ((uint32_t)(lq * 32) << 16);
      epilogue_tile<PAIR_BN, true>(p, t_addr, m0 + crank * BM + lq * 32 + lane, n0, chalf, &tmem_empty[acc], lane);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  if (warp == 1) umma::tmem_dealloc_pair(tmem_base, 2 * PAIR_BN);
}

// naive reference for the self test
__global__ void ref_gemm_kernel(const __nv_bfloat16* A, long long lda, int ta, const __nv_bfloat16* B, long long ldb,
                                int tb, int M, int N, int K, float* C) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(ta ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
    const float b = __bfloat162float(tb ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
    acc += a * b;
  }
  C[(long long)m * N + n] = acc;
}
__global__ void fill_kernel(__nv_bfloat16* p, long long n, unsigned seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = (unsigned)i * 2654435761u + seed;
  x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
  p[i] = __float2bfloat16_rn(((int)(x & 0xffff) - 32768) / 32768.f);
}
__global__ void maxdiff_kernel(const float* a, const float* b, long long n, float* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = fabsf(a[i] - b[i]);
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(d));            // d >= 0: int order == float order
}

}  // namespace
}  // namespace hnb

using namespace hnb;

// tile width: the widest of 256 / 192 / 128 columns that wastes at most ~10 % of the MMA work on the ragged last tile
// (128-wide MMAs read 8 KB of operands per 64 cycles -- the shared-memory limit -- so wider is better when it fits)
static int gemm_tile_n(int N) {
  int best = 128, best_waste = cdiv(N, 128) * 128 - N;
  const int cand[2] = {256, 192};
  for (int i = 0; i < 2; ++i) {
    const int w = cdiv(N, cand[i]) * cand[i] - N;
    if (N >= cand[i] && w * 10 <= N) return cand[i];
    if (N >= cand[i] && w < best_waste) { best = cand[i]; best_waste = w; }
  }
  return best;
}

// CTA-pair MMAs pay off when the main loop, not the epilogue, sets the pace: 256-wide tiles, at least 4 row tiles and
// at least 12 k-blocks per work item (the K = 384..512 projections drain one accumulator per 6-8 k-blocks).
static int gemm_sm_count() {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  return sms;
}

// Cost model shared by the split-K hint and the launch: waves x (k-blocks per item + fixed per-item cost), a pair's
// k-block costing ~3/4 of a single CTA's (measured on the long-K GEMMs).  Pairs only where they are not slower.
static double gemm_cost(int tiles_m, int tiles_n, int kb_per, int sk, bool pair) {
  const int sms = gemm_sm_count();
  const long long items = (long long)(pair ? cdiv(tiles_m, 2) : tiles_m) * tiles_n * sk;
  return (double)cdiv(items, pair ? sms / 2 : sms) * ((pair ? 0.75 : 1.0) * kb_per + 8.0);
}
static bool gemm_use_pair(int M, int N, int BN, int kb_per, int sk) {
  static const int cl_env = getenv("HNB_GEMM_CLUSTER") ? atoi(getenv("HNB_GEMM_CLUSTER")) : 3;
  static const int min_kb = getenv("HNB_GEMM_PAIR_MINKB") ? atoi(getenv("HNB_GEMM_PAIR_MINKB")) : 12;   // tuning knob
  if (!(cl_env == 3 && BN == 256 && cdiv(M, 128) >= 4 && kb_per >= min_kb)) return false;
  const int tm = cdiv(M, 128), tn = cdiv(N, BN);
  return gemm_cost(tm, tn, kb_per, sk, true) <= gemm_cost(tm, tn, kb_per, sk, false);
}

// Split-K factor for a weight-gradient GEMM (K = tokens): the one that minimises  waves x (k-blocks per item + fixed
// per-item cost)  for the tile grid hnb_gemm_bf16 will actually use.  The caller zero-fills C when the answer is > 1.
extern "C" int hnb_gemm_splitk_hint(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1;
  const int BN = gemm_tile_n(N), tiles_m = cdiv(M, 128), tiles_n = cdiv(N, BN), kb = cdiv(K, 64);
  int best = 1;
  double best_cost = 1e30;
  for (int sk = 1; sk <= 32 && sk * 4 <= kb; ++sk) {
    const int per = cdiv(kb, sk), eff_sk = cdiv(kb, per);
    const bool pair = gemm_use_pair(M, N, BN, per, eff_sk);         // pairs of row tiles on pairs of SMs
    const double cost = gemm_cost(tiles_m, tiles_n, per, eff_sk, pair) + 0.5 * eff_sk;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = eff_sk; }
  }
  return best;
}

// Which kernel hnb_gemm_bf16 runs for this problem: 0 = single-CTA tiles, 1 = 2-CTA multicast, 2 = CTA-pair (cta_group::2).
extern "C" int hnb_gemm_bf16_path(int M, int N, int K, int splitk) {
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  static const int cl_env = getenv("HNB_GEMM_CLUSTER") ? atoi(getenv("HNB_GEMM_CLUSTER")) : 3;
  const int BN = gemm_tile_n(N), kb_total = cdiv(K, BK);
  const int sk_req = splitk < 1 ? 1 : (splitk > kb_total ? kb_total : splitk);
  const int per_req = cdiv(kb_total, sk_req);
  if (gemm_use_pair(M, N, BN, per_req, cdiv(kb_total, per_req))) return 2;
  return (cl_env == 2 && cdiv(M, BM) >= 4 && BN != 192) ? 1 : 0;
}

extern "C" int hnb_gemm_bf16(const void* A, long long lda, int transA, const void* B, long long ldb, int transB, int M,
                             int N, int K, const float* bias, const void* R, long long ldr, void* C, long long ldc,
                             int c_dtype, int splitk, void* stream) {
  return hnb_gemm_bf16_ex(A, lda, transA, B, ldb, transB, M, N, K, bias, R, ldr, C, ldc, c_dtype, splitk, 0, 0, stream);
}

extern "C" int hnb_gemm_bf16_ex(const void* A, long long lda, int transA, const void* B, long long ldb, int transB, int M,
                                int N, int K, const float* bias, const void* R, long long ldr, void* C, long long ldc,
                                int c_dtype, int splitk, int c_col_block, long long c_block_stride, void* stream) {
  HNB_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "gemm_bf16: bad arguments");
  HNB_CHECK_ARG(c_col_block == 0 || (c_col_block > 0 && c_col_block % 32 == 0 && N % c_col_block == 0 && c_dtype == HNB_F32 && !R),
                "gemm_bf16: column-blocked C needs fp32 C, no residual, a block width that is a multiple of 32 and divides N");
  HNB_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16-byte strides)");
  HNB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "gemm_bf16: operands must be 16-byte aligned");
  HNB_CHECK_ARG(c_dtype == HNB_F32 || c_dtype == HNB_BF16, "gemm_bf16: bad output dtype");
  if (splitk < 1) splitk = 1;
  HNB_CHECK_ARG(splitk == 1 || (c_dtype == HNB_F32 && !bias && !R), "gemm_bf16: split-K needs fp32 C and no bias/residual");
  const int BN = gemm_tile_n(N);
  // HNB_GEMM_CLUSTER: 1 = single-CTA tiles, 2 = 2-CTA multicast of B (1-SM MMAs), 3 (default) = CTA-pair MMAs
  // (cta_group::2) where gemm_use_pair() says so
  static const int cl_env = getenv("HNB_GEMM_CLUSTER") ? atoi(getenv("HNB_GEMM_CLUSTER")) : 3;
  const int kb_total = cdiv(K, BK);
  const int sk_req = splitk < 1 ? 1 : (splitk > kb_total ? kb_total : splitk);
  const int per_req = cdiv(kb_total, sk_req);
  const bool pair = gemm_use_pair(M, N, BN, per_req, cdiv(kb_total, per_req));
  const int CL = (cl_env == 2 && cdiv(M, BM) >= 4 && BN != 192) ? 2 : (pair ? 2 : 1);
  CUtensorMap tmA, tmB;
  int rc;
  {
    uint64_t dims[2], st[1];
    uint32_t box[2];
    if (!transA) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)M; box[0] = BK; box[1] = BM; }
    else         { dims[0] = (uint64_t)M; dims[1] = (uint64_t)K; box[0] = 64; box[1] = BK; }
    st[0] = (uint64_t)lda * 2;
    if ((rc = make_tmap_bf16(&tmA, A, 2, dims, st, box))) return rc;
    if (!transB) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = BK; box[1] = (uint32_t)(BN / CL); }   // pair / multicast: half of B
    else         { dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64; box[1] = BK; }
    st[0] = (uint64_t)ldb * 2;
    if ((rc = make_tmap_bf16(&tmB, B, 2, dims, st, box))) return rc;
  }
  const int total_kb = cdiv(K, BK);
  if (splitk > total_kb) splitk = total_kb;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.total_kb = total_kb;
  p.kblocks_per_split = cdiv(total_kb, splitk);
  p.splitk = cdiv(total_kb, p.kblocks_per_split);
  p.tiles_m = cdiv(M, BM); p.tiles_n = cdiv(N, BN);
  p.groups_m = cdiv(p.tiles_m, CL);
  p.bias = bias; p.R = R; p.ldr = ldr; p.C = C; p.ldc = ldc;
  p.c_is_f32 = (c_dtype == HNB_F32);
  p.atomic = p.splitk > 1;
  p.cblk = c_col_block; p.cblk_stride = c_block_stride;
  const int n_items = p.groups_m * p.tiles_n * p.splitk;
  const int sms = gemm_sm_count();
  const int grid = CL * (n_items < sms / CL ? n_items : sms / CL);
  cudaStream_t st = (cudaStream_t)stream;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // common.cuh: the prologue overlaps the previous kernel's tail
  attr[1].val.programmaticStreamSerializationAllowed = hnb::pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
#define LAUNCH(TA, TB, BN_, CL_)                                                                                     \
  do {                                                                                                               \
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)gemm_bf16_kernel<TA, TB, BN_, CL_>, \
                                       GemmCfg<BN_>::SMEM));                                                         \
    cfg.dynamicSmemBytes = GemmCfg<BN_>::SMEM;                                                                       \
    HNB_CUDA_CALL(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<TA, TB, BN_, CL_>, tmA, tmB, p));                         \
  } while (0)
#define LAUNCH_BN(TA, TB)                                                                                            \
  do {                                                                                                               \
    if (BN == 256)      { if (CL == 2) LAUNCH(TA, TB, 256, 2); else LAUNCH(TA, TB, 256, 1); }                         \
    else if (BN == 192) { LAUNCH(TA, TB, 192, 1); }                                                                  \
    else                { if (CL == 2) LAUNCH(TA, TB, 128, 2); else LAUNCH(TA, TB, 128, 1); }                         \
  } while (0)
#define LAUNCH_PAIR(TA, TB)                                                                                          \
  do {                                                                                                               \
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)gemm_bf16_pair_kernel<TA, TB>, \
                                       PAIR_SMEM));                                                                  \
    cfg.dynamicSmemBytes = PAIR_SMEM;                                                                                \
    HNB_CUDA_CALL(cudaLaunchKernelEx(&cfg, gemm_bf16_pair_kernel<TA, TB>, tmA, tmB, p));                              \
  } while (0)
  if (pair) {
    if (!transA && !transB) LAUNCH_PAIR(0, 0);
    else if (!transA && transB) LAUNCH_PAIR(0, 1);
    else if (transA && !transB) LAUNCH_PAIR(1, 0);
    else LAUNCH_PAIR(1, 1);
  } else if (!transA && !transB) LAUNCH_BN(0, 0);
  else if (!transA && transB) LAUNCH_BN(0, 1);
  else if (transA && !transB) LAUNCH_BN(1, 0);
  else LAUNCH_BN(1, 1);
#undef LAUNCH_PAIR
#undef LAUNCH_BN
#undef LAUNCH
  HNB_LAUNCH_CHECK("gemm_bf16");
  return HNB_OK;
}

// Runs the four operand-major combinations (with ragged M, N, K tails) plus a split-K case against a naive kernel, once
// with 3 row tiles (single-CTA path, 128-wide tiles) and once with 5 row tiles x 256-wide tiles (the CTA-pair or
// multicast path when enabled; odd tile count: the last cluster's second CTA owns no rows).  Allocates scratch with cudaMalloc: diagnostic entry point, not a hot-path call.
extern "C" int hnb_umma_selftest(float* max_abs_err_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int K = 424, LD = 608;                                 // LD covers [M,K], [K,M], [N,K], [K,N] storage
  __nv_bfloat16 *A, *B;
  float *C, *Cref, *err;
  HNB_CUDA_CALL(cudaMalloc(&A, sizeof(__nv_bfloat16) * LD * LD));
  HNB_CUDA_CALL(cudaMalloc(&B, sizeof(__nv_bfloat16) * LD * LD));
  HNB_CUDA_CALL(cudaMalloc(&C, sizeof(float) * LD * LD));
  HNB_CUDA_CALL(cudaMalloc(&Cref, sizeof(float) * LD * LD));
  HNB_CUDA_CALL(cudaMalloc(&err, sizeof(float) * 8));
  HNB_CUDA_CALL(cudaMemsetAsync(err, 0, sizeof(float) * 8, st));
  fill_kernel<<<cdiv(LD * LD, 256), 256, 0, st>>>(A, LD * LD, 1u);
  fill_kernel<<<cdiv(LD * LD, 256), 256, 0, st>>>(B, LD * LD, 7u);
  int rc = HNB_OK;
  for (int pass = 0; pass < 2 && rc == HNB_OK; ++pass) {
    const int M = pass == 0 ? 328 : 600, N = pass == 0 ? 200 : 500;   // pass 1: 5 row tiles x 2 256-wide column tiles
    for (int combo = 0; combo < 5 && rc == HNB_OK; ++combo) {
      const int ta = combo & 1, tb = (combo >> 1) & 1;
      const int splitk = combo == 4 ? 3 : 1;
      HNB_CUDA_CALL(cudaMemsetAsync(C, 0, sizeof(float) * M * N, st));
      ref_gemm_kernel<<<dim3(cdiv(N, 128), M), 128, 0, st>>>(A, LD, ta, B, LD, tb, M, N, K, Cref);
      rc = hnb_gemm_bf16(A, LD, ta, B, LD, tb, M, N, K, nullptr, nullptr, 0, C, N, HNB_F32, splitk, stream);
      if (rc == HNB_OK) maxdiff_kernel<<<cdiv((long long)M * N, 256), 256, 0, st>>>(C, Cref, (long long)M * N, err + combo);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == HNB_OK && e != cudaSuccess) { set_error("umma_selftest: %s", cudaGetErrorString(e)); rc = HNB_ERR_CUDA; }
  if (rc == HNB_OK) cudaMemcpy(max_abs_err_host, err, sizeof(float) * 5, cudaMemcpyDeviceToHost);
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cref); cudaFree(err);
  return rc;
}
