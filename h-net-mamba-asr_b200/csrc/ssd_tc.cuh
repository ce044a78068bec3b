// Shared pieces of the tcgen05 SSD kernels (ssd_tcgen05.cu, ssd_fwd_split.cu): tile constants, the swizzled shared-memory
// addressing, bf16 pack/unpack, the per-chunk decay tables and the parameter block of the forward kernels.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace hnb {
namespace {

constexpr int TQ = 128;                 // chunk length
constexpr int TP = 64;                  // head dim
constexpr int TN = 128;                 // state dim
constexpr int FWD_THREADS = 512;          // forward: 16 warps
constexpr int BWD_THREADS = 512;          // dx / dbc
constexpr int HALF = TQ * 128;          // bytes of one [128 rows x 128 B] swizzled block (16 KB)
constexpr int TAB_FLOATS = 9 * TQ + 8;  // cs, dt, w, ecs, eq [128 each] | f[4][128] | warp totals
constexpr int TAB_BYTES = TAB_FLOATS * 4;
constexpr int EXTRA_FLOATS = 12 * TQ + 3 * BWD_THREADS + 8;   // dx kernel: per-column-group / per-thread partial sums

__device__ __forceinline__ uint32_t swz(int row, int chunk16) {          // byte offset inside a [rows x 128 B] SW128 block
  return (uint32_t)row * 128u + (uint32_t)((chunk16 ^ (row & 7)) << 4);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& r, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Per-chunk tables.  Called by the first `nt` threads of the CTA (every warp but the MMA-issuing one, so that the
// tables of the next step are built WHILE the issuer feeds the tensor core); synchronises them on named barrier 1.
//   cs[t]  inclusive cumsum of dt*A      dt[t]  (0 beyond the row's length)
//   w[t]   e^{cs_last - cs_t} dt_t       ecs[t] e^{cs_t}        eq[t] e^{cs_last - cs_t}
//   f[I][s] = e^{cs_{32I-1} - cs_s}  for s < 32 I  (block factor of the decay matrix), else 0
__device__ __forceinline__ void build_tables(const float* __restrict__ dtp, int H, int qv, float A, float* tab, int nt) {
  float* s_cs = tab; float* s_dt = tab + TQ; float* s_w = tab + 2 * TQ; float* s_ecs = tab + 3 * TQ;
  float* s_eq = tab + 4 * TQ; float* s_f = tab + 5 * TQ; float* s_tot = tab + 9 * TQ;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < TQ) {
    const float d = (tid < qv) ? __ldg(dtp + (long long)tid * H) : 0.f;
    s_dt[tid] = d;
    float v = d * A;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    s_cs[tid] = v;
    if (lane == 31) s_tot[tid >> 5] = v;
  }
  named_sync(1, nt);
  if (tid < TQ) {
    float add = 0.f;
    for (int w = 0; w < (tid >> 5); ++w) add += s_tot[w];
    s_cs[tid] += add;
  }
  named_sync(1, nt);
  const float cs_last = s_cs[TQ - 1];
  if (tid < TQ) {
    const float cs = s_cs[tid];
    const float eq = __expf(cs_last - cs);
    s_eq[tid] = eq;
    s_w[tid] = eq * s_dt[tid];
    s_ecs[tid] = __expf(cs);
  }
  for (int i = tid; i < 4 * TQ; i += nt) {
    const int I = i >> 7, s = i & (TQ - 1);
    s_f[i] = (I > 0 && s < 32 * I) ? __expf(s_cs[32 * I - 1] - s_cs[s]) : 0.f;
  }
  named_sync(1, nt);
}

// Balanced split of the lower-triangular 128 x 128 score tile over the 16 epilogue warps: warp (lq, cg) owns, in
// each of the four 32-column blocks k, the 8 columns 32k + 8cg .. +7 of rows 32lq .. +31.  Every warp of a row
// block then carries the same load (k < lq: one multiply per element through the block factor table; k == lq: one
// exp per element; k > lq: zeros) instead of one warp doing the whole exp-heavy diagonal block while others idle.
// l[j] = L[t, s0 + j], j < 8, for column block k (warp-uniform case split).
__device__ __forceinline__ void decay8(float* l, int t, int I, int k, int s0, float cs_t, float e_ref, const float* tab) {
  if (k < I) {
    const float4 a = *reinterpret_cast<const float4*>(tab + 5 * TQ + I * TQ + s0);
    const float4 b = *reinterpret_cast<const float4*>(tab + 5 * TQ + I * TQ + s0 + 4);
    l[0] = e_ref * a.x; l[1] = e_ref * a.y; l[2] = e_ref * a.z; l[3] = e_ref * a.w;
    l[4] = e_ref * b.x; l[5] = e_ref * b.y; l[6] = e_ref * b.z; l[7] = e_ref * b.w;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(tab + s0);
    const float4 b = *reinterpret_cast<const float4*>(tab + s0 + 4);
    const float c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = (s0 + j <= t) ? __expf(cs_t - c[j]) : 0.f;
  }
}
__device__ __forceinline__ void load8(float* v, const float* src) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

struct FwdParams {
  const float* dt;        // [ndir*B*L, H]
  const float* A_log;     // [ndir, H]
  const float* Dskip;     // [ndir, H]
  __nv_bfloat16* y;       // [ndir*B*L, di]
  __nv_bfloat16* states;  // [ndir*B, H, nc, 128(n), 64(p)]  state ENTERING each chunk
  const float* tables;    // [ndir*B, H, nc, TAB_FLOATS]
  const __nv_bfloat16* xconv;   // [ndir*B*L, di + 2N]  (the two-CTA kernel reads x for the D x term from here)
  int* counter;                 // two-CTA kernel: next unclaimed (row, head) item minus gridDim.x (zeroed by ssd_tables_kernel)
  int ndirB, B, L, H, di, nc;
  FastDiv dH, dB, dnc;
  long long* dbg;         // optional [8] phase-cycle accumulators of CTA 0 (HNB_SSD_DEBUG=1)
};

}  // namespace
}  // namespace hnb
