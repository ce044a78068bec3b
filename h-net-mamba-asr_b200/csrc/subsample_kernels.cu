// ConvSubsampling4 front end (reference src/dcasr/models/encoder.py:55-70): the first Conv2d(1 -> C, k3, s2) + ReLU,
// forward and backward.  The layer has ONE input channel: 9 MACs per output, but a [B, C, T1, F1] result of
// 0.96 GB (bf16, 40 x 16 s).  Through cuDNN + ATen the step paid for that tensor eight times (fp32->bf16 casts, a
// separate ReLU forward/backward, NCHW<->NHWC transposes, an fp32 SGEMM-style conv); here it is written once (forward,
// straight into the NHWC layout cuDNN's tensor-core kernels of the second convolution want) and read once (backward:
// ReLU mask from the saved output, dW1 / db1 reduced on the fly).
//
// Work decomposition: a block owns RB consecutive (b, t) output rows; a thread owns CPT consecutive channels of one
// of PX pixel lanes and keeps its 9*CPT taps in registers; the three input rows a (b, t) row needs are staged in shared
// memory.  NHWC stores/loads are 16 bytes (forward, 8 channels) / 8 bytes (backward, 4 channels) per thread, contiguous
// across the threads of a pixel.
#include <cstdlib>

#include "common.cuh"

namespace hnb {

constexpr int SUB_RB = 8;            // (b, t) rows per block

// PK (F even): the nine taps of a pixel are consumed as PAIRS on the packed fp32 pipe of sm_100 (FFMA2): (tap 0, tap 1),
// (3, 4), (6, 7) are aligned 8-byte shared-memory loads, (2, 5) is assembled from two 4-byte loads, tap 8 stays scalar: 4 FFMA2 +
// 1 FFMA + 1 FADD and 6 loads per output instead of 9 FFMA and 9 loads, with the same registers (the channel-pair form tried
// before, DESIGN.md section 9, needed duplicated inputs and lost the occupancy).  Both kernels are issue-bound at their
// occupancy (ncu: issue slots 71-74 % busy, FMA pipe ~35 %); measured, the pairs pay in the backward (559 -> 507 us) and not
// in the forward (354 -> 397 us), so only the backward uses them by default.
template <int CPT, bool PK, typename TO = __nv_bfloat16>
__global__ void __launch_bounds__(256, PK ? 4 : 1)
sub_conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int B, int T,
                     int F, int C, int T1, int F1, TO* __restrict__ out) {
  pdl_enter();
  extern __shared__ __align__(16) float s_in[];         // [3][F]
  const int CG = C / CPT;                               // threads per pixel
  const int cgp = threadIdx.x % CG, pl = threadIdx.x / CG, PX = blockDim.x / CG;
  const int c0 = cgp * CPT;
  float wt[CPT][9], bs[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int k = 0; k < 9; ++k) wt[i][k] = __ldg(w + (long long)(c0 + i) * 9 + k);
    bs[i] = __ldg(bias + c0 + i);
  }
  const long long rows = (long long)B * T1;
  for (long long r = (long long)blockIdx.x * SUB_RB; r < min(rows, (long long)(blockIdx.x + 1) * SUB_RB); ++r) {
    const int b = (int)(r / T1), t = (int)(r % T1);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) s_in[i] = x[((long long)b * T + 2 * t + i / F) * F + i % F];
    __syncthreads();
    for (int f = pl; f < F1; f += PX) {
      float o[CPT];
      if constexpr (PK) {
        const float2 p0 = *reinterpret_cast<const float2*>(s_in + 2 * f), p1 = *reinterpret_cast<const float2*>(s_in + F + 2 * f),
                     p2 = *reinterpret_cast<const float2*>(s_in + 2 * F + 2 * f);
        const float2 q = make_float2(s_in[2 * f + 2], s_in[F + 2 * f + 2]);
        const float x8 = s_in[2 * F + 2 * f + 2];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float2 a = __ffma2_rn(make_float2(wt[i][0], wt[i][1]), p0, make_float2(bs[i], 0.f));
          a = __ffma2_rn(make_float2(wt[i][3], wt[i][4]), p1, a);
          a = __ffma2_rn(make_float2(wt[i][6], wt[i][7]), p2, a);
          a = __ffma2_rn(make_float2(wt[i][2], wt[i][5]), q, a);
          o[i] = fmaxf(fmaf(wt[i][8], x8, a.x) + a.y, 0.f);
        }
      } else {
        float xin[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) xin[k] = s_in[(k / 3) * F + 2 * f + k % 3];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          float a = bs[i];
#pragma unroll
          for (int k = 0; k < 9; ++k) a = fmaf(wt[i][k], xin[k], a);
          o[i] = fmaxf(a, 0.f);
        }
      }
      stv<TO, CPT>(out + ((r * F1 + f) * C + c0), o);
    }
  }
}

// backward: dA1 and the saved output A1 (both NHWC, bf16) -> dW1 [C, 9], db1 [C] (fp32, accumulated).  The ReLU mask
// is A1 > 0, as in the reference's threshold_backward; the input needs no gradient.  A block reduces its rows in
// registers, then across its pixel lanes through shared memory, then issues 10 vector reductions per channel quad.
template <int CPT, int NT, int MINB, bool PK>
__global__ void __launch_bounds__(NT, MINB)
sub_conv1_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ a1, const __nv_bfloat16* __restrict__ dout,
                     int B, int T, int F, int C, int T1, int F1, int rows_per_block, float* __restrict__ dw,
                     float* __restrict__ db) {
  pdl_enter();
  static_assert(CPT == 4, "vector reductions below assume channel quads");
  extern __shared__ __align__(16) float s_in[];         // [3][F], then reused for the block reduction
  const int CG = C / CPT;
  const int cgp = threadIdx.x % CG, pl = threadIdx.x / CG, PX = blockDim.x / CG;
  const int c0 = cgp * CPT;
  float gw[CPT][9], gb[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int k = 0; k < 9; ++k) gw[i][k] = 0.f;
    gb[i] = 0.f;
  }
  const long long rows = (long long)B * T1;
  const long long r_beg = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(rows, r_beg + rows_per_block);
  // Four pixel vectors of each tensor per thread and step, and the NEXT step's eight loads are issued before the current
  // step's arithmetic (also across the row boundary): without that every step was one exposed DRAM round trip
  // (ncu: long-scoreboard stalls 4.4 per issue, 3.1 TB/s).
  const int NG = (F1 + 4 * PX - 1) / (4 * PX);
  uint2 gn[4], an[4];
  auto issue = [&](long long r, int g) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = pl + (4 * g + u) * PX;
      if (f < F1) {
        gn[u] = __ldg(reinterpret_cast<const uint2*>(dout + ((r * F1 + f) * C + c0)));
        an[u] = __ldg(reinterpret_cast<const uint2*>(a1 + ((r * F1 + f) * C + c0)));
      }
    }
  };
  if (r_beg < r_end) issue(r_beg, 0);
  for (long long r = r_beg; r < r_end; ++r) {
    const int b = (int)(r / T1), t = (int)(r % T1);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * F; i += blockDim.x) s_in[i] = x[((long long)b * T + 2 * t + i / F) * F + i % F];
    __syncthreads();
    for (int g = 0; g < NG; ++g) {
      uint2 gr[4], ar[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { gr[u] = gn[u]; ar[u] = an[u]; }
      if (g + 1 < NG) issue(r, g + 1);
      else if (r + 1 < r_end) issue(r + 1, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pl + (4 * g + u) * PX;
        if (f >= F1) break;
        float gq[CPT], a[CPT];
        {
          const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gr[u]);
          const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&ar[u]);
          const float2 g0 = __bfloat1622float2(gh[0]), g1 = __bfloat1622float2(gh[1]);
          const float2 a0 = __bfloat1622float2(ah[0]), a1v = __bfloat1622float2(ah[1]);
          gq[0] = g0.x; gq[1] = g0.y; gq[2] = g1.x; gq[3] = g1.y;
          a[0] = a0.x; a[1] = a0.y; a[2] = a1v.x; a[3] = a1v.y;
        }
        if constexpr (PK) {                               // tap pairs on the packed fp32 pipe (see the forward)
          const float2 p0 = *reinterpret_cast<const float2*>(s_in + 2 * f), p1 = *reinterpret_cast<const float2*>(s_in + F + 2 * f),
                       p2 = *reinterpret_cast<const float2*>(s_in + 2 * F + 2 * f);
          const float2 q = make_float2(s_in[2 * f + 2], s_in[F + 2 * f + 2]);
          const float x8 = s_in[2 * F + 2 * f + 2];
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const float gi = a[i] > 0.f ? gq[i] : 0.f;
            const float2 g2 = make_float2(gi, gi);
            gb[i] += gi;
            float2 t;
            t = __ffma2_rn(g2, p0, make_float2(gw[i][0], gw[i][1])); gw[i][0] = t.x; gw[i][1] = t.y;
            t = __ffma2_rn(g2, p1, make_float2(gw[i][3], gw[i][4])); gw[i][3] = t.x; gw[i][4] = t.y;
            t = __ffma2_rn(g2, p2, make_float2(gw[i][6], gw[i][7])); gw[i][6] = t.x; gw[i][7] = t.y;
            t = __ffma2_rn(g2, q, make_float2(gw[i][2], gw[i][5])); gw[i][2] = t.x; gw[i][5] = t.y;
            gw[i][8] = fmaf(gi, x8, gw[i][8]);
          }
        } else {
          float xin[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) xin[k] = s_in[(k / 3) * F + 2 * f + k % 3];
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const float gi = a[i] > 0.f ? gq[i] : 0.f;
            gb[i] += gi;
#pragma unroll
            for (int k = 0; k < 9; ++k) gw[i][k] = fmaf(gi, xin[k], gw[i][k]);
          }
        }
      }
    }
  }
  __syncthreads();
  float* red = s_in;                                     // [PX-1][CG][40]  (the launch sizes shared memory for it)
  if (pl > 0) {
    float* dst = red + ((long long)(pl - 1) * CG + cgp) * 40;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
#pragma unroll
      for (int k = 0; k < 9; ++k) dst[i * 9 + k] = gw[i][k];
      dst[36 + i] = gb[i];
    }
  }
  __syncthreads();
  if (pl == 0) {
    for (int q = 0; q < PX - 1; ++q) {
      const float* src = red + ((long long)q * CG + cgp) * 40;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
#pragma unroll
        for (int k = 0; k < 9; ++k) gw[i][k] += src[i * 9 + k];
        gb[i] += src[36 + i];
      }
    }
    // dw rows of a channel quad are 36 contiguous floats starting at a multiple of 144 bytes: 9 + 1 vector reductions
    float* wq = dw + (long long)c0 * 9;
    const float* flat = &gw[0][0];
#pragma unroll
    for (int k = 0; k < 9; ++k)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wq + 4 * k), "f"(flat[4 * k]), "f"(flat[4 * k + 1]),
                   "f"(flat[4 * k + 2]), "f"(flat[4 * k + 3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(db + c0), "f"(gb[0]), "f"(gb[1]), "f"(gb[2]), "f"(gb[3])
                 : "memory");
  }
}

// ---- bias + ReLU over an NHWC tensor viewed as [rows, C] (the second convolution's output) ----------------------
// forward: in place.  backward: dpre = dout * [out > 0] and db[c] += sum_rows dpre in the same pass (ATen ran a strided
// broadcast add, a clamp, a threshold_backward and a 116 M-element bf16 column reduction: four passes each way).
// fp32 twin (decode: the second convolution's fp32 NHWC output), 4 channels = 16 bytes per thread
__global__ void __launch_bounds__(256)
bias_relu_fwd_f32_kernel(float* __restrict__ x, const float* __restrict__ bias, long long n4, int C4) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % C4));
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

__global__ void __launch_bounds__(256)
bias_relu_fwd_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ bias, long long rows, int C) {
  pdl_enter();
  const int CG = C / 8, cgp = threadIdx.x % CG, rl = threadIdx.x / CG, RL = blockDim.x / CG;
  float b[8];
  ldv<float, 8>(bias + cgp * 8, b);
  for (long long r = (long long)blockIdx.x * RL + rl; r < rows; r += (long long)gridDim.x * RL) {
    float v[8];
    __nv_bfloat16* p = x + r * C + cgp * 8;
    ldv<__nv_bfloat16, 8>(p, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i] + b[i], 0.f);
    stv<__nv_bfloat16, 8>(p, v);
  }
}

__global__ void __launch_bounds__(256)
bias_relu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                     __nv_bfloat16* __restrict__ dpre, float* __restrict__ db, long long rows, int C) {
  pdl_enter();
  extern __shared__ float s_red[];                      // [RL-1][CG][8]
  const int CG = C / 8, cgp = threadIdx.x % CG, rl = threadIdx.x / CG, RL = blockDim.x / CG;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r = (long long)blockIdx.x * RL + rl; r < rows; r += (long long)gridDim.x * RL) {
    float g[8], o[8];
    const long long off = r * C + cgp * 8;
    ldv<__nv_bfloat16, 8>(dout + off, g);
    ldv<__nv_bfloat16, 8>(out + off, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = o[i] > 0.f ? g[i] : 0.f; acc[i] += g[i]; }
    stv<__nv_bfloat16, 8>(dpre + off, g);
  }
  if (rl > 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[((rl - 1) * CG + cgp) * 8 + i] = acc[i];
  }
  __syncthreads();
  if (rl == 0) {
    for (int q = 0; q < RL - 1; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += s_red[(q * CG + cgp) * 8 + i];
    float* d = db + cgp * 8;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
  }
}

}  // namespace hnb

using namespace hnb;

static int sub_check(const char* who, int B, int T, int F, int C) {
  if (!(B > 0 && T >= 3 && F >= 3 && C > 0)) { set_error("%s: bad sizes", who); return HNB_ERR_INVALID_ARG; }
  if (C % 8 || C > 1024) {
    set_error("%s: channels must be a multiple of 8, at most 1024 (got %d)", who, C);
    return HNB_ERR_UNSUPPORTED;
  }
  return HNB_OK;
}

extern "C" int hnb_subsample_conv1_fwd_dt(const float* feats, const float* w, const float* bias, int B, int T, int F, int C,
                                          void* out, int out_dtype, void* stream) {
  HNB_CHECK_ARG(out_dtype == HNB_BF16 || out_dtype == HNB_F32, "subsample_conv1_fwd: output dtype must be bf16 or fp32");
  if (out_dtype == HNB_BF16) return hnb_subsample_conv1_fwd(feats, w, bias, B, T, F, C, out, stream);
  HNB_CHECK_ARG(feats && w && bias && out, "subsample_conv1_fwd: null pointer");
  int rc = sub_check("subsample_conv1_fwd", B, T, F, C);
  if (rc) return rc;
  const int T1 = (T - 3) / 2 + 1, F1 = (F - 3) / 2 + 1;
  const long long rows = (long long)B * T1;
  const int CG = C / 4, PX = 256 / CG > 0 ? 256 / CG : 1;   // fp32 output (decode): 4 channels = one 16-byte store per thread and pixel
  hnb::launch_pdl(sub_conv1_fwd_kernel<4, false, float>, dim3(cdiv(rows, SUB_RB)), dim3(CG * PX), 3 * F * sizeof(float), (cudaStream_t)stream,
      feats, w, bias, B, T, F, C, T1, F1, (float*)out);
  HNB_LAUNCH_CHECK("subsample_conv1_fwd");
  return HNB_OK;
}

extern "C" int hnb_subsample_conv1_fwd(const float* feats, const float* w, const float* bias, int B, int T, int F, int C,
                                       void* out, void* stream) {
  HNB_CHECK_ARG(feats && w && bias && out, "subsample_conv1_fwd: null pointer");
  int rc = sub_check("subsample_conv1_fwd", B, T, F, C);
  if (rc) return rc;
  const int T1 = (T - 3) / 2 + 1, F1 = (F - 3) / 2 + 1;
  const long long rows = (long long)B * T1;
  // packed tap pairs: measured SLOWER in the forward (397 vs 354 us: the kernel is not FMA-issue-bound once a store follows every 20
  // instructions), faster in the backward (507 vs 559 us); off here unless HNB_SUB_FWD_PACKED=1
  static const bool pk_env = [] { const char* e = getenv("HNB_SUB_FWD_PACKED"); return e && atoi(e) == 1; }();
  const bool pk = pk_env && F % 2 == 0;                  // packed tap pairs need 8-byte aligned input pairs
  static const int cpt = [] { const char* e = getenv("HNB_SUB_FWD_CPT"); return e && atoi(e) == 8 ? 8 : 4; }();
  if (cpt == 4) {                                        // half the taps per thread, 64 registers, five resident blocks: 401 -> 334 us (HNB_SUB_FWD_CPT=8: the former kernel)
    const int CG = C / 4, PX = 256 / CG > 0 ? 256 / CG : 1;
    if (pk) hnb::launch_pdl(sub_conv1_fwd_kernel<4, true>, dim3(cdiv(rows, SUB_RB)), dim3(CG * PX), 3 * F * sizeof(float), (cudaStream_t)stream,
        feats, w, bias, B, T, F, C, T1, F1, (__nv_bfloat16*)out);
    else hnb::launch_pdl(sub_conv1_fwd_kernel<4, false>, dim3(cdiv(rows, SUB_RB)), dim3(CG * PX), 3 * F * sizeof(float), (cudaStream_t)stream,
        feats, w, bias, B, T, F, C, T1, F1, (__nv_bfloat16*)out);
  } else {
    const int CG = C / 8, PX = 256 / CG > 0 ? 256 / CG : 1;
    hnb::launch_pdl(sub_conv1_fwd_kernel<8, false>, dim3(cdiv(rows, SUB_RB)), dim3(CG * PX), 3 * F * sizeof(float), (cudaStream_t)stream,
        feats, w, bias, B, T, F, C, T1, F1, (__nv_bfloat16*)out);
  }
  HNB_LAUNCH_CHECK("subsample_conv1_fwd");
  return HNB_OK;
}

extern "C" int hnb_subsample_conv1_bwd(const float* feats, const void* a1, const void* dout, int B, int T, int F, int C,
                                       float* dw, float* db, void* stream) {
  HNB_CHECK_ARG(feats && a1 && dout && dw && db, "subsample_conv1_bwd: null pointer");
  HNB_CHECK_ARG(((reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(db)) & 15) == 0,
                "subsample_conv1_bwd: dw and db must be 16-byte aligned");
  int rc = sub_check("subsample_conv1_bwd", B, T, F, C);
  if (rc) return rc;
  const int T1 = (T - 3) / 2 + 1, F1 = (F - 3) / 2 + 1;
  const long long rows = (long long)B * T1;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  static const int nt = [] { const char* e = getenv("HNB_SUB_BWD_THREADS"); return e && atoi(e) == 384 ? 384 : 256; }();   // 384-thread blocks measured slower (658 vs 626 us)
  static const bool pk_env = [] { const char* e = getenv("HNB_SUB_PACKED"); return !(e && atoi(e) == 0); }();
  const bool pk = pk_env && F % 2 == 0;
  static const int minb = [] { const char* e = getenv("HNB_SUB_BWD_MINB"); return e && atoi(e) == 2 ? 2 : 3; }();   // 80 registers (12 bytes spilled), four resident 192-thread blocks: 556 us; 114 registers, no spill: 621 us
  const int CG = C / 4, PX = nt / CG > 0 ? nt / CG : 1;
  static const int bpsm = [] { const char* e = getenv("HNB_SUB_BWD_BLOCKS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 12; }();   // measured 3 / 4 / 6 / 8 / 12 / 16 per SM: 705 / 617 / 626 / 596 / 585 / 597 us
  int blocks = sms * bpsm;                               // whole waves of the resident blocks; each block ends in 10 reductions per quad
  if (blocks > rows) blocks = (int)rows;
  const int rpb = cdiv(rows, blocks);
  blocks = cdiv(rows, rpb);
  size_t smem = 3 * (size_t)F * sizeof(float);
  const size_t red = (size_t)(PX > 1 ? PX - 1 : 0) * CG * 40 * sizeof(float);
  if (red > smem) smem = red;
  if (nt == 384 && CG * PX <= 384) {                     // two resident blocks of up to 384 threads: a third more bytes in flight per SM
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)sub_conv1_bwd_kernel<4, 384, 2, false>, (int)smem));
    hnb::launch_pdl(sub_conv1_bwd_kernel<4, 384, 2, false>, dim3(blocks), dim3(CG * PX), smem, (cudaStream_t)stream, feats, (const __nv_bfloat16*)a1,
        (const __nv_bfloat16*)dout, B, T, F, C, T1, F1, rpb, dw, db);
  } else if (minb != 2 && pk) {
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)sub_conv1_bwd_kernel<4, 256, 3, true>, (int)smem));
    hnb::launch_pdl(sub_conv1_bwd_kernel<4, 256, 3, true>, dim3(blocks), dim3(CG * PX), smem, (cudaStream_t)stream, feats, (const __nv_bfloat16*)a1,
        (const __nv_bfloat16*)dout, B, T, F, C, T1, F1, rpb, dw, db);
  } else if (minb != 2) {
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)sub_conv1_bwd_kernel<4, 256, 3, false>, (int)smem));
    hnb::launch_pdl(sub_conv1_bwd_kernel<4, 256, 3, false>, dim3(blocks), dim3(CG * PX), smem, (cudaStream_t)stream, feats, (const __nv_bfloat16*)a1,
        (const __nv_bfloat16*)dout, B, T, F, C, T1, F1, rpb, dw, db);
  } else {
    HNB_CUDA_CALL(hnb_set_max_smem((const void*)sub_conv1_bwd_kernel<4, 256, 2, false>, (int)smem));
    hnb::launch_pdl(sub_conv1_bwd_kernel<4, 256, 2, false>, dim3(blocks), dim3(CG * PX), smem, (cudaStream_t)stream, feats, (const __nv_bfloat16*)a1,
        (const __nv_bfloat16*)dout, B, T, F, C, T1, F1, rpb, dw, db);
  }
  HNB_LAUNCH_CHECK("subsample_conv1_bwd");
  return HNB_OK;
}

static int bias_relu_geometry(const char* who, long long rows, int C, int* threads, int* blocks) {
  if (!(rows > 0 && C > 0 && C % 8 == 0 && C <= 2048)) { set_error("%s: C must be a multiple of 8, at most 2048", who); return HNB_ERR_INVALID_ARG; }
  const int CG = C / 8, RL = 256 / CG > 0 ? 256 / CG : 1;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  *threads = CG * RL;
  long long b = (rows + RL - 1) / RL;
  *blocks = (int)(b < (long long)sms * 8 ? b : (long long)sms * 8);
  return HNB_OK;
}

extern "C" int hnb_bias_relu_fwd(void* x, const float* bias, long long rows, int C, void* stream) {
  HNB_CHECK_ARG(x && bias && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "bias_relu_fwd: null or misaligned pointer");
  int threads, blocks, rc = bias_relu_geometry("bias_relu_fwd", rows, C, &threads, &blocks);
  if (rc) return rc;
  hnb::launch_pdl(bias_relu_fwd_kernel, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, (__nv_bfloat16*)x, bias, rows, C);
  HNB_LAUNCH_CHECK("bias_relu_fwd");
  return HNB_OK;
}

extern "C" int hnb_bias_relu_fwd_f32(float* x, const float* bias, long long rows, int C, void* stream) {
  HNB_CHECK_ARG(x && bias && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "bias_relu_fwd_f32: null or misaligned pointer");
  HNB_CHECK_ARG(rows > 0 && C > 0 && C % 4 == 0, "bias_relu_fwd_f32: C must be a multiple of 4");
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const long long n4 = rows * (C / 4);
  const long long want = (n4 + 255) / 256;
  const int blocks = (int)(want < (long long)sms * 16 ? want : (long long)sms * 16);
  hnb::launch_pdl(bias_relu_fwd_f32_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, bias, n4, C / 4);
  HNB_LAUNCH_CHECK("bias_relu_fwd_f32");
  return HNB_OK;
}

extern "C" int hnb_bias_relu_bwd(const void* dout, const void* out, void* dpre, float* db, long long rows, int C, void* stream) {
  HNB_CHECK_ARG(dout && out && dpre && db && (reinterpret_cast<uintptr_t>(db) & 15) == 0, "bias_relu_bwd: null or misaligned pointer");
  int threads, blocks, rc = bias_relu_geometry("bias_relu_bwd", rows, C, &threads, &blocks);
  if (rc) return rc;
  const int CG = C / 8, RL = threads / CG;
  const size_t smem = (size_t)(RL > 1 ? RL - 1 : 0) * CG * 8 * sizeof(float);
  hnb::launch_pdl(bias_relu_bwd_kernel, dim3(blocks), dim3(threads), smem, (cudaStream_t)stream, (const __nv_bfloat16*)dout, (const __nv_bfloat16*)out,
                                                                       (__nv_bfloat16*)dpre, db, rows, C);
  HNB_LAUNCH_CHECK("bias_relu_bwd");
  return HNB_OK;
}
