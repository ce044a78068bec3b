// One Mamba block (pre-norm, bidirectional or not) per host call: the composites below run the per-kernel entry points of
// this library back to back on one stream, inside one workspace whose layout they own.  Nothing new is computed here;
// what changes is the host cost: a block used to be ~8 (forward) + ~12 (backward) Python -> ctypes -> C round trips with
// a dozen tensor allocations, ~13 ms of host time per encoder step of 20 blocks -- more than half of the 20 ms the GPU
// needs, and the first thing to break 8-GPU scaling on a host whose cores the ranks share.
//
//   forward : LayerNorm -> weight packing -> in-projection (both directions, one GEMM) -> conv1d + SiLU / softplus
//             -> SSD scan -> gated RMSNorm -> out-projection + residual          (src/dcasr/models/mamba_block.py:50-56)
//   backward: the same chain in reverse, weight gradients into a caller-zeroed arena.
#include "common.cuh"

namespace hnb {
namespace {

constexpr size_t ALIGN = 256;
inline size_t up(size_t v) { return (v + ALIGN - 1) / ALIGN * ALIGN; }
inline size_t esz(int dt) { return dt == HNB_BF16 ? 2 : 4; }

struct Dims {
  int B, L, d, ndir, di, N, H, act;
  long long T() const { return (long long)B * L; }
  int C() const { return di + 2 * N; }
  int dip() const { return 2 * di + 2 * N + H; }
  int dstride() const { return (dip() + 7) / 8 * 8; }
  int ldz() const { return ndir * dstride(); }
};

// forward workspace (kept for the backward)
enum { F_H2, F_MEAN, F_RSTDLN, F_WIN, F_WOUT, F_SMALL, F_ZX, F_XCONV, F_DT, F_Y, F_SSD, F_YN, F_RSTD, F_GEMM, F_COUNT };
// fp32 activations: the large projections run as bf16-piece GEMMs on the tensor cores (hnb_gemm_f32_tc) and need scratch
inline bool f32_tc(long long M, long long N, long long K) { return (double)M * (double)N * (double)K >= 67108864.0; }
size_t f32_gemm_scratch(int act, long long M, long long N, long long K) {
  return (act == HNB_F32 && f32_tc(M, N, K)) ? (size_t)hnb_gemm_f32_tc_ws_bytes((int)M, (int)N, (int)K) : 0;
}
inline size_t max2(size_t x, size_t y) { return x > y ? x : y; }

size_t fwd_layout(const Dims& m, size_t* off) {
  const size_t a = esz(m.act), T = (size_t)m.T();
  const size_t gs = max2(f32_gemm_scratch(m.act, m.T(), m.ldz(), m.d), f32_gemm_scratch(m.act, m.T(), m.d, m.ndir * m.di));
  const size_t sz[F_COUNT] = {
      T * m.d * a, T * 4, T * 4, (size_t)m.ldz() * m.d * a, (size_t)m.d * m.ndir * m.di * a,
      (size_t)m.ndir * (m.C() * 5 + 3 * m.H + m.di) * 4, T * m.ldz() * a, (size_t)m.ndir * T * m.C() * a,
      (size_t)m.ndir * T * m.H * 4, (size_t)m.ndir * T * m.di * a,
      (size_t)hnb_ssd_ws_bytes(m.ndir, m.B, m.L, m.di, m.N, m.H), T * m.ndir * m.di * a, (size_t)m.ndir * T * 4, gs};
  size_t o = 0;
  for (int i = 0; i < F_COUNT; ++i) { off[i] = o; o += up(sz[i]); }
  return o;
}
// the packed fp32 vectors inside F_SMALL, in the order hnb_pack_mixer_params* writes its stacks
struct Small { float *conv_w, *conv_b, *dt_bias, *A_log, *D, *norm_w; };
Small small_ptrs(const Dims& m, uint8_t* base) {
  float* p = reinterpret_cast<float*>(base);
  Small s;
  s.conv_w = p; p += (size_t)m.ndir * m.C() * 4;
  s.conv_b = p; p += (size_t)m.ndir * m.C();
  s.dt_bias = p; p += (size_t)m.ndir * m.H;
  s.A_log = p; p += (size_t)m.ndir * m.H;
  s.D = p; p += (size_t)m.ndir * m.H;
  s.norm_w = p;
  return s;
}

// weights of one block as hnb_pack_mixer_stack leaves them: Win | Wout | small vectors
size_t packed_layout(const Dims& m, size_t* off3) {
  const size_t a = esz(m.act);
  const size_t sz[3] = {(size_t)m.ldz() * m.d * a, (size_t)m.d * m.ndir * m.di * a,
                        (size_t)m.ndir * (m.C() * 5 + 3 * m.H + m.di) * 4};
  size_t o = 0;
  for (int i = 0; i < 3; ++i) { off3[i] = o; o += up(sz[i]); }
  return o;
}

// backward scratch (dead when the call returns)
enum { S_DA, S_DYN, S_DZX, S_DY, S_DXC, S_DBC, S_DDT, S_WS2, S_DH2, S_GEMM, S_COUNT };
size_t bwd_layout(const Dims& m, int parts, size_t* off) {
  const size_t a = esz(m.act), T = (size_t)m.T();
  const long long nd = (long long)m.ndir * m.di;
  const size_t gs = max2(max2(f32_gemm_scratch(m.act, m.T(), nd, m.d), f32_gemm_scratch(m.act, m.d, nd, m.T())),
                         max2(f32_gemm_scratch(m.act, m.T(), m.d, m.ldz()), f32_gemm_scratch(m.act, m.ldz(), m.d, m.T())));
  const size_t sz[S_COUNT] = {
      T * m.d * a, T * m.ndir * m.di * a, T * m.ldz() * a, (size_t)m.ndir * T * m.di * a, (size_t)m.ndir * T * m.di * a,
      (size_t)parts * m.ndir * T * 2 * m.N * a, (size_t)m.ndir * T * m.H * 4,
      (size_t)hnb_ssd_ws_bytes(m.ndir, m.B, m.L, m.di, m.N, m.H), T * m.d * a, gs};
  size_t o = 0;
  for (int i = 0; i < S_COUNT; ++i) { off[i] = o; o += up(sz[i]); }
  return o;
}

// gradient arena (fp32): dWout [ndir][d, di] | dWin [ndir*dstride, d] | conv_w | conv_b | norm_w |
// dA_log | dD | ddt_bias | LayerNorm dgamma, dbeta
enum { A_WOUT, A_WIN, A_CW, A_CB, A_NW, A_DA, A_DD, A_DTB, A_LN, A_COUNT };
size_t arena_layout(const Dims& m, size_t* off) {      // in floats
  const size_t sz[A_COUNT] = {(size_t)m.d * m.ndir * m.di, (size_t)m.ldz() * m.d, (size_t)m.ndir * m.C() * 4,
                              (size_t)m.ndir * m.C(), (size_t)m.ndir * m.di, (size_t)m.ndir * m.H, (size_t)m.ndir * m.H,
                              (size_t)m.ndir * m.H, (size_t)2 * m.d};
  size_t o = 0;
  for (int i = 0; i < A_COUNT; ++i) { off[i] = o; o += sz[i]; }
  return o;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
    h[0] = __floats2bfloat162_rn(v.x, v.y); h[1] = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

int check_dims(const char* who, const Dims& m, int x_dtype) {
  HNB_CHECK_ARG(m.B > 0 && m.L > 0 && m.d > 0 && (m.ndir == 1 || m.ndir == 2) && m.di > 0 && m.N > 0 && m.H > 0,
                "%s: bad dimensions", who);
  HNB_CHECK_ARG(m.act == HNB_BF16 || m.act == HNB_F32, "%s: bad activation dtype", who);
  HNB_CHECK_ARG(x_dtype == HNB_BF16 || x_dtype == HNB_F32, "%s: bad residual dtype", who);
  HNB_CHECK_ARG(!(m.act == HNB_F32 && x_dtype != HNB_F32), "%s: fp32 activations need an fp32 residual stream", who);
  HNB_CHECK_ARG(m.d % 4 == 0 && m.di % 4 == 0, "%s: d_model and d_inner must be multiples of 4", who);
  return HNB_OK;
}

#define HNB_TRY(expr) do { int rc__ = (expr); if (rc__ != HNB_OK) return rc__; } while (0)

// C = op(A) op(B) in the activation dtype's GEMM (tensor cores for bf16, exact CUDA-core GEMM for fp32)
int gemm(int act, const void* A, long long lda, int tA, const void* Bm, long long ldb, int tB, int M, int N, int K,
         const void* R, long long ldr, void* Cm, long long ldc, int c_dtype, int splitk, void* st, void* scratch = nullptr) {
  if (act == HNB_BF16) return hnb_gemm_bf16(A, lda, tA, Bm, ldb, tB, M, N, K, nullptr, R, ldr, Cm, ldc, c_dtype, splitk, st);
  if (scratch && f32_tc(M, N, K))
    return hnb_gemm_f32_tc((const float*)A, lda, tA, (const float*)Bm, ldb, tB, M, N, K, nullptr, (const float*)R, ldr,
                           (float*)Cm, ldc, scratch, st);
  return hnb_gemm_f32((const float*)A, lda, tA, (const float*)Bm, ldb, tB, M, N, K, nullptr, (const float*)R, ldr,
                      (float*)Cm, ldc, 0, st);
}

}  // namespace
}  // namespace hnb

using namespace hnb;

extern "C" long long hnb_block_packed_layout(int d, int ndir, int di, int N, int H, int act_dtype, long long* off3) {
  const Dims m{1, 1, d, ndir, di, N, H, act_dtype};
  size_t o[3];
  const size_t n = packed_layout(m, o);
  if (off3) for (int i = 0; i < 3; ++i) off3[i] = (long long)o[i];
  return (long long)n;
}
extern "C" long long hnb_block_packed_bytes(int d, int ndir, int di, int N, int H, int act_dtype) {
  return hnb_block_packed_layout(d, ndir, di, N, H, act_dtype, nullptr);
}

extern "C" long long hnb_block_fwd_ws_bytes(int B, int L, int d, int ndir, int di, int N, int H, int act_dtype) {
  const Dims m{B, L, d, ndir, di, N, H, act_dtype};
  size_t off[F_COUNT];
  return (long long)fwd_layout(m, off);
}
extern "C" long long hnb_block_bwd_ws_bytes(int B, int L, int d, int ndir, int di, int N, int H, int act_dtype, int ssd_impl) {
  const Dims m{B, L, d, ndir, di, N, H, act_dtype};
  size_t off[S_COUNT];
  return (long long)bwd_layout(m, hnb_ssd_dbc_parts(ndir, B, L, H, ssd_impl), off);
}
extern "C" long long hnb_block_grad_floats(int B, int L, int d, int ndir, int di, int N, int H, long long* offsets) {
  const Dims m{B, L, d, ndir, di, N, H, HNB_BF16};
  size_t off[A_COUNT];
  const size_t n = arena_layout(m, off);
  if (offsets) for (int i = 0; i < A_COUNT; ++i) offsets[i] = (long long)off[i];
  return (long long)n;
}

extern "C" int hnb_block_fwd(const void* x, int x_dtype, const int32_t* lengths, const float* ln_w, const float* ln_b,
                             const void* const* params, int B, int L, int d, int ndir, int di, int N, int H,
                             int act_dtype, int ssd_impl, const void* packed, void* out, void* ws, void* stream) {
  const Dims m{B, L, d, ndir, di, N, H, act_dtype};
  HNB_TRY(check_dims("block_fwd", m, x_dtype));
  HNB_CHECK_ARG(x && ln_w && ln_b && (params || packed) && out && ws, "block_fwd: null pointer");
  if (!packed)
    for (int i = 0; i < 8 * ndir; ++i) HNB_CHECK_ARG(params[i] != nullptr, "block_fwd: null parameter pointer");
  size_t off[F_COUNT];
  fwd_layout(m, off);
  uint8_t* w = static_cast<uint8_t*>(ws);
  const long long T = m.T();
  const int dstride = m.dstride(), ldz = m.ldz();
  void* h2 = w + off[F_H2];
  float* mean = reinterpret_cast<float*>(w + off[F_MEAN]);
  float* rstd_ln = reinterpret_cast<float*>(w + off[F_RSTDLN]);
  size_t po[3];
  packed_layout(m, po);
  uint8_t* pk = const_cast<uint8_t*>(static_cast<const uint8_t*>(packed));     // weights packed once per stack, or here
  void* Win = packed ? pk + po[0] : w + off[F_WIN]; void* Wout = packed ? pk + po[1] : w + off[F_WOUT];
  const Small s = small_ptrs(m, packed ? pk + po[2] : w + off[F_SMALL]);
  void* zx = w + off[F_ZX]; void* xconv = w + off[F_XCONV];
  float* dt = reinterpret_cast<float*>(w + off[F_DT]);
  void* y = w + off[F_Y]; void* ssd = w + off[F_SSD]; void* yn = w + off[F_YN];
  float* rstd = reinterpret_cast<float*>(w + off[F_RSTD]);
  const float* const* P = reinterpret_cast<const float* const*>(params);   // per direction: in_w, conv_w, conv_b, dt_bias, A_log, D, norm_w, out_w

  HNB_TRY(hnb_layernorm_fwd(x, x_dtype, ln_w, ln_b, T, d, 1e-5f, h2, act_dtype, mean, rstd_ln, stream));
  if (packed) {
    // nothing to pack
  } else if (ndir == 2)
    HNB_TRY(hnb_pack_mixer_params2(P[0], P[7], P[1], P[2], P[3], P[4], P[5], P[6], P[8], P[15], P[9], P[10], P[11], P[12],
                                   P[13], P[14], d, di, N, H, dstride, Win, Wout, act_dtype, s.conv_w, s.conv_b, s.dt_bias,
                                   s.A_log, s.D, s.norm_w, stream));
  else
    HNB_TRY(hnb_pack_mixer_params(P[0], P[7], P[1], P[2], P[3], P[4], P[5], P[6], 0, 1, d, di, N, H, dstride, Win, Wout,
                                  act_dtype, s.conv_w, s.conv_b, s.dt_bias, s.A_log, s.D, s.norm_w, stream));
  void* gsc = act_dtype == HNB_F32 ? w + off[F_GEMM] : nullptr;
  HNB_TRY(gemm(act_dtype, h2, d, 0, Win, d, 0, (int)T, ldz, d, nullptr, 0, zx, ldz, act_dtype, 1, stream, gsc));
  HNB_TRY(hnb_conv_fwd(zx, act_dtype, ldz, dstride, lengths, s.conv_w, s.conv_b, s.dt_bias, ndir, B, L, di, N, H, xconv, dt,
                       stream));
  HNB_TRY(hnb_ssd_fwd(xconv, act_dtype, dt, s.A_log, s.D, ndir, B, L, di, N, H, y, ssd, ssd_impl, stream));
  HNB_TRY(hnb_gated_norm_fwd(y, zx, act_dtype, ldz, dstride, lengths, s.norm_w, ndir, B, L, di, 1e-5f, yn, rstd, stream));
  // out-projection of both directions (K = ndir*di) with the residual in its epilogue; the residual stream keeps its dtype
  HNB_TRY(gemm(act_dtype, yn, (long long)ndir * di, 0, Wout, (long long)ndir * di, 0, (int)T, d, ndir * di, x, d, out, d,
               x_dtype, 1, stream, gsc));
  return HNB_OK;
}

extern "C" int hnb_block_bwd(const void* dout, const void* x, int x_dtype, const int32_t* lengths, const float* ln_w,
                             const void* ws, int B, int L, int d, int ndir, int di, int N, int H, int act_dtype,
                             int ssd_impl, const void* packed, void* dx, float* grads, int zero_grads, void* scratch,
                             void* stream) {
  const Dims m{B, L, d, ndir, di, N, H, act_dtype};
  HNB_TRY(check_dims("block_bwd", m, x_dtype));
  HNB_CHECK_ARG(dout && x && ln_w && ws && dx && grads && scratch, "block_bwd: null pointer");
  size_t off[F_COUNT], so[S_COUNT], ao[A_COUNT];
  fwd_layout(m, off);
  const int parts = hnb_ssd_dbc_parts(ndir, B, L, H, ssd_impl);
  bwd_layout(m, parts, so);
  arena_layout(m, ao);
  const uint8_t* w = static_cast<const uint8_t*>(ws);
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  const long long T = m.T();
  const int dstride = m.dstride(), ldz = m.ldz(), dip = m.dip();
  const void* h2 = w + off[F_H2];
  const float* mean = reinterpret_cast<const float*>(w + off[F_MEAN]);
  const float* rstd_ln = reinterpret_cast<const float*>(w + off[F_RSTDLN]);
  size_t po[3];
  packed_layout(m, po);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  const void* Win = packed ? pk + po[0] : w + off[F_WIN]; const void* Wout = packed ? pk + po[1] : w + off[F_WOUT];
  const Small s = small_ptrs(m, const_cast<uint8_t*>(packed ? pk + po[2] : w + off[F_SMALL]));
  const void* zx = w + off[F_ZX]; const void* xconv = w + off[F_XCONV];
  const float* dt = reinterpret_cast<const float*>(w + off[F_DT]);
  const void* y = w + off[F_Y]; const void* ssd = w + off[F_SSD]; const void* yn = w + off[F_YN];
  const float* rstd = reinterpret_cast<const float*>(w + off[F_RSTD]);
  void* dyn = sc + so[S_DYN]; void* dzx = sc + so[S_DZX]; void* dy = sc + so[S_DY]; void* dxc = sc + so[S_DXC];
  void* dBC = sc + so[S_DBC]; float* ddt = reinterpret_cast<float*>(sc + so[S_DDT]); void* ws2 = sc + so[S_WS2];
  void* dh2 = sc + so[S_DH2];
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_grads) HNB_CUDA_CALL(cudaMemsetAsync(grads, 0, arena_layout(m, ao) * sizeof(float), st));

  // gradient of the block output in the activation dtype (the residual stream may be fp32 under bf16 autocast)
  const void* da = dout;
  if (x_dtype != act_dtype) {
    void* buf = sc + so[S_DA];
    const long long n = T * d;
    const long long want = (n / 4 + 255) / 256;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    if (x_dtype == HNB_F32) hnb::launch_pdl(cast_f32_bf16_kernel, dim3(grid), dim3(256), 0, st, (const float*)dout, (__nv_bfloat16*)buf, n / 4);
    else hnb::launch_pdl(cast_bf16_f32_kernel, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dout, (float*)buf, n);
    HNB_LAUNCH_CHECK("block_bwd cast");
    da = buf;
  }
  const int bf = act_dtype == HNB_BF16;
  void* gsc = act_dtype == HNB_F32 ? sc + so[S_GEMM] : nullptr;
  HNB_TRY(gemm(act_dtype, da, d, 0, Wout, (long long)ndir * di, 1, (int)T, ndir * di, d, nullptr, 0, dyn, (long long)ndir * di,
               act_dtype, 1, stream, gsc));                                                   // d ynorm
  // dWout of both directions from ONE GEMM, each direction's [d, di] matrix contiguous (column-blocked C): the parameter's
  // .grad can then alias the arena instead of being a strided copy of it
  if (bf && (ndir == 1 || di % 32 == 0)) {
    HNB_TRY(hnb_gemm_bf16_ex(da, d, 1, yn, (long long)ndir * di, 1, d, ndir * di, (int)T, nullptr, nullptr, 0, grads + ao[A_WOUT],
                             di, HNB_F32, hnb_gemm_splitk_hint(d, ndir * di, (int)T), ndir > 1 ? di : 0, (long long)d * di,
                             stream));
  } else {
    for (int r = 0; r < ndir; ++r) {
      const size_t a = esz(act_dtype);
      const void* yn_r = static_cast<const uint8_t*>(yn) + (size_t)r * di * a;
      HNB_TRY(gemm(act_dtype, da, d, 1, yn_r, (long long)ndir * di, 1, d, di, (int)T, nullptr, 0,
                   grads + ao[A_WOUT] + (long long)r * d * di, di, HNB_F32, bf ? hnb_gemm_splitk_hint(d, di, (int)T) : 1, stream, gsc));
    }
  }
  if (dstride != dip) {                                       // pad columns of dzxbcdt feed the two GEMMs below
    const size_t a = esz(act_dtype);
    HNB_CUDA_CALL(cudaMemset2DAsync(static_cast<uint8_t*>(dzx) + (size_t)dip * a, (size_t)dstride * a, 0,
                                    (size_t)(dstride - dip) * a, (size_t)T * ndir, st));
  }
  HNB_TRY(hnb_gated_norm_bwd(dyn, y, zx, act_dtype, ldz, dstride, lengths, s.norm_w, rstd, ndir, B, L, di, dy, dzx,
                             grads + ao[A_NW], stream));
  HNB_TRY(hnb_ssd_bwd(dy, xconv, y, act_dtype, dt, s.A_log, s.D, ssd, ndir, B, L, di, N, H, dxc, dBC, parts, ddt,
                      grads + ao[A_DA], grads + ao[A_DD], ws2, ssd_impl, stream));
  HNB_TRY(hnb_conv_bwd(zx, dxc, act_dtype, ldz, dstride, dBC, ddt, lengths, s.conv_w, s.conv_b, s.dt_bias, ndir, B, L, di, N,
                       H, dzx, grads + ao[A_CW], grads + ao[A_CB], grads + ao[A_DTB], parts, stream));
  HNB_TRY(gemm(act_dtype, dzx, ldz, 0, Win, d, 1, (int)T, d, ldz, nullptr, 0, dh2, d, act_dtype, 1, stream, gsc));   // d LN output
  HNB_TRY(gemm(act_dtype, dzx, ldz, 1, h2, d, 1, ldz, d, (int)T, nullptr, 0, grads + ao[A_WIN], d, HNB_F32,
               bf ? hnb_gemm_splitk_hint(ldz, d, (int)T) : 1, stream, gsc));                                         // dWin
  HNB_TRY(hnb_layernorm_bwd(dh2, act_dtype, x, x_dtype, ln_w, mean, rstd_ln, dout, T, d, dx, x_dtype, grads + ao[A_LN],
                            grads + ao[A_LN] + d, stream));
  return HNB_OK;
}
