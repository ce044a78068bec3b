/* hnet_b200 — C ABI of the B200-native H-Net / Mamba-2 encoder hot path.
 *
 * The reference (anshulk-cmu/H-Net-Mamba-ASR, package `dcasr`) has no FFI of its own: the
 * boundary is Python class identity.  This header is the seam a maintainer
 * binds with ctypes; every entry point names the reference code whose
 * arithmetic it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - no allocation, no internal streams or threads: the caller owns every buffer and workspace;
 *  - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns at once;
 *  - return 0 on success, a HNB_ERR_* code otherwise; hnb_last_error() gives the message
 *    (thread-local);
 *  - activations are row-major [B, L, D] flattened to [B*L, D]; `dtype` says how they are stored,
 *    arithmetic is always fp32 (tensor-core contractions: bf16 operands, fp32 accumulate);
 *  - "scan order" (Mamba kernels): direction 0 is natural time; direction 1 walks each row's
 *    valid span [0, len_b) back to front and leaves the right padding in place, which is the
 *    reference's reverse_sequences() (src/dcasr/models/mamba_block.py:19-28) done by index
 *    arithmetic instead of two gathers.
 */
#ifndef HNET_B200_H
#define HNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNB_OK 0
#define HNB_ERR_INVALID_ARG 1
#define HNB_ERR_CUDA 2
#define HNB_ERR_UNSUPPORTED 3

#define HNB_F32 0
#define HNB_BF16 1

/* ---- library ---------------------------------------------------------------------------- */
int hnb_version(void);
const char* hnb_last_error(void);
/* number of kernels this library has launched since the last reset (bench.py: gpu_launches) */
long long hnb_launch_count(void);
void hnb_reset_launch_count(void);

/* ---- H-Net dynamic chunking stage -------------------------------------------------------- */

/* RoutingModule.forward epilogue (src/dcasr/models/hnet_chunk.py:98-108) fused with the
 * ratio-loss partial sums (hnet_chunk.py:126-134).
 * qk      [B*L, ldqk]  q in columns [0,D), k in columns [D,2D)   (the W_q/W_k GEMM output)
 * mask    [B*L] uint8 (1 = valid) or NULL
 * p, b    [B*L] out, stored as pb_dtype: p = 0.5(1-cos(q_t,k_{t-1})), p[:,0]=1, clamp, b=[p>=.5], *mask
 * partial [nblk*4] float workspace, nblk = hnb_router_num_partials(B*L) */
int hnb_router_num_partials(long long n_tokens);
int hnb_router_fwd(const void* qk, int qk_dtype, long long ldqk, const uint8_t* mask,
                   int B, int L, int D, float eps, void* p, void* b, int pb_dtype,
                   float* partial, void* stream);

/* ratio_loss (hnet_chunk.py:117-136) + kept_fraction (hnet_chunk.py:193-194) from the partial sums.
 * stats[8] out: [0]=ratio loss, [1]=kept fraction, [2]=F, [3]=G, [4]=denominator, [5]=sum b, [6]=sum p */
int hnb_ratio_finalize(const float* partial, int nblk, float N, float* stats, void* stream);

/* backward of the two above: d qk from d p.
 * dp_ext  [B*L] float or NULL : gradient arriving at p from other consumers (STE, EMA, user code)
 * dratio  [1] float or NULL   : gradient arriving at the ratio-loss scalar */
int hnb_router_bwd(const void* qk, int qk_dtype, long long ldqk, const uint8_t* mask,
                   int B, int L, int D, float eps, const float* dp_ext, const float* dratio,
                   const float* stats, float N, void* dqk, void* stream);

/* membership = clamp(cumsum(b>0.5)-1, 0) (int64) and per-row counts (hnet_chunk.py:182-184) as ONE
 * single-pass segmented decoupled-look-back scan over the flattened [B*L] flags.
 * ws: workspace of hnb_boundary_scan_ws_bytes(B*L) bytes (zeroed by the call). */
long long hnb_boundary_scan_ws_bytes(long long n_tokens);
int hnb_boundary_scan(const void* b, int pb_dtype, int B, int L, int64_t* membership,
                      int32_t* counts, void* ws, void* stream);

/* downsample stream compaction (hnet_chunk.py:188-192 and :214-218): kept rows of x -> z[B,M,D]
 * (pad slots zero-filled), z_mask, downsampled P, and starts[b,j] = frame index of the j-th kept frame
 * (L for pad slots).  M = max(1, max counts) is read by the host between scan and compaction. */
int hnb_compact_rows(const void* x, int x_dtype, const void* p, const void* b, int pb_dtype,
                     const int64_t* membership, const int32_t* counts, int B, int L, int D, int M,
                     void* z, uint8_t* z_mask, float* P, int32_t* starts, void* stream);

/* backward of the compaction: dx[b,t] (+)= b>0.5 ? dz[b, membership[b,t]] : 0 */
int hnb_compact_rows_bwd(const void* dz, int dtype, const void* b, int pb_dtype,
                         const int64_t* membership, int B, int L, int D, int M,
                         void* dx, int accumulate, void* stream);

/* DynamicChunker._ema (hnet_chunk.py:226-248) as a linear-time scan:
 * out_0 = x_0, out_t = pc_t x_t + (1-pc_t) out_{t-1}, pc = clamp(P, p_clamp, 1-p_clamp) */
int hnb_ema_fwd(const void* x, int dtype, const float* P, int B, int M, int D, float p_clamp,
                void* out, void* stream);
/* backward: dx, dP (dP must be zero-initialised; zero gradient where P is outside the clamp band) */
int hnb_ema_bwd(const void* dout, const void* x, const void* out, int dtype, const float* P,
                int B, int M, int D, float p_clamp, void* dx, float* dP, void* stream);

/* upsample gather + confidence STE (+ the encoder's residual add) (hnet_chunk.py:220-224,
 * encoder.py:130):  y[b,t] = resid[b,t] + zbar[b, membership[b,t]] * (c + (1-c)),  c = b ? p : 1-p */
int hnb_upsample_fwd(const void* zbar, int z_dtype, const int64_t* membership, const void* p,
                     const void* b, int pb_dtype, const void* resid, int B, int L, int D, int M,
                     void* y, int y_dtype, void* stream);
/* backward: dzbar[b,j] = sum over the frames of chunk j of dy * ste;  dp[b,t] = +-<dy[b,t], zbar[b,j]> */
int hnb_upsample_bwd(const void* dy, int y_dtype, const void* zbar, int z_dtype,
                     const int64_t* membership, const int32_t* starts, const int32_t* counts,
                     const void* p, const void* b, int pb_dtype, int B, int L, int D, int M,
                     void* dzbar, float* dp, void* stream);

/* masked sums for a stand-alone ratio_loss(p, b, N, mask) call: partial[nblk*4] as in hnb_router_fwd */
int hnb_masked_sums(const void* p, const void* b, int pb_dtype, const uint8_t* mask, long long n,
                    float* partial, void* stream);

/* ---- FixedPoolChunker (src/dcasr/models/fixed_pool.py:54-106): the fixed-stride control of the learned chunker ----
 * Window of frame t: w(t) = min(t / stride, M-1) (the reference clamps a padded tail into the last window, :80-83).
 * reduce: z[b,j] = sum over the frames of window j of m[b,t] x[b,t], divided by max(cnt[b,j],1) when `normalize`
 *   (chunk(): masked mean, :84-89; normalize = 0 and mask = NULL give the backward of dechunk()).
 *   mask [B,L] uint8 or NULL (all frames); cnt [B,M] float (valid frames per window) or NULL. */
int hnb_window_reduce(const void* x, int x_dtype, const uint8_t* mask, int B, int L, int D, int M, int stride,
                      int normalize, void* z, int z_dtype, float* cnt, void* stream);
/* broadcast: out[b,t] = m[b,t] / max(cnt[b,w],1) * z[b, w(t)] (+ resid[b,t])
 *   (dechunk(): gather, :96-104, with mask = cnt = NULL; with both it is the backward of the masked mean). */
int hnb_window_broadcast(const void* z, int z_dtype, const uint8_t* mask, const float* cnt, const void* resid,
                         int B, int L, int D, int M, int stride, void* out, int out_dtype, void* stream);

/* ---- Mamba-2 block ----------------------------------------------------------------------- */

/* nn.LayerNorm (src/dcasr/models/mamba_block.py:51,73). mean/rstd [rows] float are saved for backward. */
int hnb_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta,
                      long long rows, int d, float eps, void* y, int y_dtype,
                      float* mean, float* rstd, void* stream);
/* dx = (dres ? dres : 0) + LN'(dy);  dgamma/dbeta [d] are ACCUMULATED (caller zero-initialises) */
int hnb_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma,
                      const float* mean, const float* rstd, const void* dres, long long rows, int d,
                      void* dx, int dx_dtype, float* dgamma, float* dbeta, void* stream);

/* Layout of the fused in-projection output `zxbcdt` [B*L, ldz] used by the kernels below, for ndir
 * directions, d_inner = di, d_state = N, heads = H, conv channels C = di + 2N.  Direction r owns the
 * column block [r*dstride, r*dstride + 2di+2N+H) in mamba_ssm's own in_proj order  z | xBC | dt :
 *   z   : [r*dstride, +di)      xBC : [r*dstride + di, +C)      dt : [r*dstride + di + C, +H)
 * dstride >= 2di+2N+H is rounded up so that every block starts 16-byte aligned; the pad columns come out
 * of the GEMM as zeros (zero weight rows) and are never read. */

/* causal depthwise conv1d(k=4)+SiLU over xBC and dt = softplus(dt_raw + dt_bias), both written in
 * SCAN ORDER (mamba_ssm causal_conv1d_fn + _chunk_cumsum_fwd's softplus).
 * conv_w [ndir, C, 4], conv_b [ndir, C], dt_bias [ndir, H] float; lengths [B] int32 or NULL.
 * xconv [ndir, B*L, C] (act dtype), dt [ndir, B*L, H] float. */
int hnb_conv_fwd(const void* zxbcdt, int dtype, long long ldz, long long dstride, const int32_t* lengths,
                 const float* conv_w, const float* conv_b, const float* dt_bias,
                 int ndir, int B, int L, int di, int N, int H,
                 void* xconv, float* dt, void* stream);
/* backward: reads dxc [ndir,B*L,di] and dBC [dbc_parts][ndir,B*L,2N] (both act dtype: d wrt conv'd x | B | C; the
 * dbc_parts partial sums of hnb_ssd_bwd are added up on the fly), ddt [ndir,B*L,H] float (all scan order); writes the
 * xBC and dt columns of dzxbcdt (natural order) and ACCUMULATES dconv_w, dconv_b, ddt_bias. */
int hnb_conv_bwd(const void* zxbcdt, const void* dxc, int dtype, long long ldz, long long dstride, const void* dBC,
                 const float* ddt, const int32_t* lengths, const float* conv_w, const float* conv_b,
                 const float* dt_bias, int ndir, int B, int L, int di, int N, int H,
                 void* dzxbcdt, float* dconv_w, float* dconv_b, float* ddt_bias, int dbc_parts, void* stream);

/* SSD selective scan in scan order (mamba_ssm mamba_chunk_scan_combined):
 *   h_t = exp(dt_t A) h_{t-1} + dt_t B_t (x) x_t ,  y_t = C_t . h_t + D x_t
 * xconv [ndir,B*L,C] holds x | B | C; dt [ndir,B*L,H]; A_log, Dskip [ndir,H] float.
 * y [ndir,B*L,di] (act dtype).  states: workspace of hnb_ssd_ws_bytes() bytes (per-chunk states,
 * kept for the backward).  impl: 0 = CUDA-core fp32 (exact, any dtype), 1 = tcgen05 (bf16 only; forward kernel chosen by
 * shape between 4 and 5; backward: state-gradient pass + ONE fused dx | dB/dC kernel), 2 = tcgen05 with the
 * one-CTA-per-SM forward kernel (forward only; same outputs), 3 = tcgen05 with the three-kernel backward of round 1
 * (state gradients, dx, dB/dC; same outputs, kept under test as the yard-stick of the fused kernel), 4 = as 1 with the
 * persistent forward kernel per (row, head), two CTAs per SM, 5 = as 1 with the split forward (chunk-state pass, then a
 * scan over every chunk of every row at once with the score tile shared by the heads; the faster one on long rows). */
long long hnb_ssd_ws_bytes(int ndir, int B, int L, int di, int N, int H);
int hnb_ssd_chunk(void);
int hnb_ssd_fwd(const void* xconv, int dtype, const float* dt, const float* A_log, const float* Dskip,
                int ndir, int B, int L, int di, int N, int H, void* y, void* states, int impl,
                void* stream);
/* backward.  dy [ndir,B*L,di].  Outputs: dxc [ndir,B*L,di] (act dtype), dBC [dbc_parts][ndir,B*L,2N] (act dtype),
 * ddt [ndir,B*L,H] float, and ACCUMULATED dA_log, dD [ndir,H].
 * dbc_parts = hnb_ssd_dbc_parts(...): the tcgen05 dB/dC kernel may split the heads into two groups (1 or 2 parts) so that its
 * work items fill whole waves of SMs; each group writes its own partial sum and hnb_conv_bwd adds the parts up.
 * ws2: second workspace of hnb_ssd_ws_bytes() bytes. */
int hnb_ssd_dbc_parts(int ndir, int B, int L, int H, int impl);
/* host-only (tests): out[0..G] = the cut positions of that schedule for G SMs, as indices into the sequence of
 * (row-chunk item, head) steps (items: every full chunk first, then the partly filled last chunks) */
int hnb_ssd_span_cuts(int ndirB, int L, int H, int G, int* out);
int hnb_ssd_bwd(const void* dy, const void* xconv, const void* y, int dtype, const float* dt,
                const float* A_log, const float* Dskip, const void* states,
                int ndir, int B, int L, int di, int N, int H,
                void* dxc, void* dBC, int dbc_parts, float* ddt, float* dA_log, float* dD, void* ws2, int impl,
                void* stream);

/* gated RMSNorm  rmsnorm(y * silu(z)) * w  (mamba_ssm RMSNormGated, norm_before_gate=False, eps 1e-5)
 * reading y in scan order and writing natural order: out [B*L, ndir*di] (the out-projection operand).
 * norm_w [ndir, di] float; rstd [ndir, B*L] float saved for backward. */
int hnb_gated_norm_fwd(const void* y, const void* zxbcdt, int dtype, long long ldz, long long dstride,
                       const int32_t* lengths, const float* norm_w, int ndir, int B, int L, int di,
                       float eps, void* out, float* rstd, void* stream);
/* backward: dout [B*L, ndir*di] -> dy [ndir,B*L,di] (scan order), z columns of dzxbcdt, ACCUMULATED dnorm_w */
int hnb_gated_norm_bwd(const void* dout, const void* y, const void* zxbcdt, int dtype, long long ldz,
                       long long dstride, const int32_t* lengths, const float* norm_w, const float* rstd,
                       int ndir, int B, int L, int di, void* dy, void* dzxbcdt, float* dnorm_w,
                       void* stream);

/* per-step packing of ONE direction's mixer parameters (fp32 masters, mamba_ssm layout) into the fused operands:
 *   in_proj.weight [2di+2N+H, d] -> rows [dir*dstride, ...) of Win [ndir*dstride, d] (pad rows zeroed)
 *   out_proj.weight [d, di]      -> columns [dir*di, +di) of Wout [d, ndir*di]          (both cast to w_dtype)
 *   conv1d.weight [C,4], conv1d.bias [C], dt_bias/A_log/D [H], norm.weight [di] -> slot `dir` of their [ndir,...] stacks */
int hnb_pack_mixer_params(const float* in_w, const float* out_w, const float* conv_w, const float* conv_b,
                          const float* dt_bias, const float* A_log, const float* Dk, const float* norm_w,
                          int dir, int ndir, int d, int di, int N, int H, int dstride, void* Win, void* Wout,
                          int w_dtype, float* conv_w_o, float* conv_b_o, float* dt_bias_o, float* A_log_o,
                          float* D_o, float* norm_w_o, void* stream);
/* the same for BOTH directions of a bidirectional block in one launch (ndir = 2; *0 = forward, *1 = backward mixer) */
int hnb_pack_mixer_params2(const float* in_w0, const float* out_w0, const float* conv_w0, const float* conv_b0,
                           const float* dt_bias0, const float* A_log0, const float* Dk0, const float* norm_w0,
                           const float* in_w1, const float* out_w1, const float* conv_w1, const float* conv_b1,
                           const float* dt_bias1, const float* A_log1, const float* Dk1, const float* norm_w1,
                           int d, int di, int N, int H, int dstride, void* Win, void* Wout, int w_dtype,
                           float* conv_w_o, float* conv_b_o, float* dt_bias_o, float* A_log_o, float* D_o,
                           float* norm_w_o, void* stream);

/* ---- one Mamba block per host call ----------------------------------------------------------
 * MambaBlock.forward / its backward (src/dcasr/models/mamba_block.py:50-56) as ONE call each: the composites run the
 * entry points above back to back on `stream` inside one workspace whose layout the library owns.  Same arithmetic,
 * same kernels; what they save is host time (a block was ~20 Python -> C round trips and a dozen allocations).
 *   x [B*L, d] (x_dtype: the residual stream, fp32 or bf16), ln_w / ln_b [d] float, lengths [B] int32 or NULL,
 *   params: HOST array of 8*ndir device pointers (fp32 masters), per direction in mamba_ssm order
 *           in_proj.w, conv1d.w, conv1d.b, dt_bias, A_log, D, norm.w, out_proj.w
 *   act_dtype: dtype of the activations / GEMM operands (bf16 under autocast); ssd_impl as in hnb_ssd_fwd
 *   packed: NULL (the call casts / stacks the block's weights itself into ws), or this block's slice of the buffer
 *           hnb_pack_mixer_stack filled for the whole stack in one launch (then `params` may be NULL); the backward
 *           gets the same pointer.
 *   out [B*L, d] (x_dtype) = x + mixers(LayerNorm(x));  ws: hnb_block_fwd_ws_bytes() bytes, kept for the backward. */
long long hnb_block_fwd_ws_bytes(int B, int L, int d, int ndir, int di, int N, int H, int act_dtype);
long long hnb_block_packed_bytes(int d, int ndir, int di, int N, int H, int act_dtype);
long long hnb_block_packed_layout(int d, int ndir, int di, int N, int H, int act_dtype, long long* off3);
/* params: HOST array [nlayers][ndir][8] of device pointers (order as for hnb_block_fwd); packed: nlayers *
 * hnb_block_packed_bytes() bytes.  One launch for the whole MambaStack (src/dcasr/models/mamba_block.py:59-73). */
int hnb_pack_mixer_stack(const void* const* params, int nlayers, int ndir, int d, int di, int N, int H, int act_dtype,
                         void* packed, void* stream);
int hnb_block_fwd(const void* x, int x_dtype, const int32_t* lengths, const float* ln_w, const float* ln_b,
                  const void* const* params, int B, int L, int d, int ndir, int di, int N, int H,
                  int act_dtype, int ssd_impl, const void* packed, void* out, void* ws, void* stream);
/* backward: dout [B*L, d] (x_dtype) -> dx [B*L, d] (x_dtype) and the parameter gradients, written into the fp32 arena
 * `grads` of hnb_block_grad_floats() floats (zero_grads != 0: the call clears the arena first on `stream`; 0: it
 * ACCUMULATES into what the caller left there).  offsets[9] (in floats) of that arena:
 * dWout [ndir][d, di] (one contiguous matrix per direction) | dWin [ndir*dstride, d] (direction r = rows [r*dstride, +2di+2N+H)) | dconv_w [ndir,C,4] |
 * dconv_b [ndir,C] | dnorm_w [ndir,di] | dA_log | dD | ddt_bias [ndir,H] | LayerNorm dgamma, dbeta [2,d].
 * scratch: hnb_block_bwd_ws_bytes() bytes, dead when the call returns. */
long long hnb_block_bwd_ws_bytes(int B, int L, int d, int ndir, int di, int N, int H, int act_dtype, int ssd_impl);
long long hnb_block_grad_floats(int B, int L, int d, int ndir, int di, int N, int H, long long* offsets);
int hnb_block_bwd(const void* dout, const void* x, int x_dtype, const int32_t* lengths, const float* ln_w,
                  const void* ws, int B, int L, int d, int ndir, int di, int N, int H, int act_dtype,
                  int ssd_impl, const void* packed, void* dx, float* grads, int zero_grads, void* scratch, void* stream);

/* ---- dense projections (in_proj / out_proj / router W_q,W_k / proj_in,out) ---------------- */
/* C[M,N] = op(A) op(B) (+ bias[N]) (+ R[M,N]) on tcgen05 tensor cores, bf16 operands, fp32 accumulate.
 *   transA = 0: A is [M,K] row-major (lda >= K);  1: A is [K,M] row-major (lda >= M)
 *   transB = 0: B is [N,K] row-major (ldb >= K), i.e. C = A B^T (nn.Linear);  1: B is [K,N] row-major
 *   c_dtype: HNB_BF16 or HNB_F32;  bias float or NULL;  R (c_dtype, ldr) or NULL
 *   splitk > 1: K is split and fp32 partial tiles are atomically added into C (C must be fp32 and
 *   pre-initialised, e.g. zero) — used for weight gradients where K = tokens. */
int hnb_gemm_bf16(const void* A, long long lda, int transA, const void* B, long long ldb, int transB,
                  int M, int N, int K, const float* bias, const void* R, long long ldr,
                  void* C, long long ldc, int c_dtype, int splitk, void* stream);

/* The same with a column-blocked C: c_col_block > 0 stores C as N / c_col_block separate row-major [M, c_col_block]
 * matrices (row stride ldc), c_block_stride elements apart -- the out-projection weight gradient of BOTH directions comes
 * out of one GEMM as two contiguous [d, d_inner] matrices (mamba_ssm out_proj.weight layout, mamba_block.py:45-47).
 * fp32 C only, no residual, c_col_block % 32 == 0. */
int hnb_gemm_bf16_ex(const void* A, long long lda, int transA, const void* B, long long ldb, int transB,
                     int M, int N, int K, const float* bias, const void* R, long long ldr,
                     void* C, long long ldc, int c_dtype, int splitk, int c_col_block, long long c_block_stride,
                     void* stream);

/* Recommended split-K factor for hnb_gemm_bf16 on this (M, N, K): fills whole waves of the tile grid the library will
 * use.  > 1 means: zero-fill an fp32 C and pass the value as `splitk`.  (Weight gradients of the reference's
 * nn.Linear layers, src/dcasr/models/mamba_block.py:45-47 via mamba_ssm; K = tokens.) */
int hnb_gemm_splitk_hint(int M, int N, int K);
/* Which kernel hnb_gemm_bf16 picks for (M, N, K, splitk): 0 = single-CTA tiles, 1 = 2-CTA multicast of B,
 * 2 = CTA-pair tiles (tcgen05.mma.cta_group::2).  Test / diagnosis aid: the choice is a cost model, not an argument. */
int hnb_gemm_bf16_path(int M, int N, int K, int splitk);
/* exact fp32 GEMM on CUDA cores (decode / fp32 parity path), same operand conventions */
int hnb_gemm_f32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                 int M, int N, int K, const float* bias, const float* R, long long ldr,
                 float* C, long long ldc, int accumulate, void* stream);
/* the same product on the tcgen05 tensor cores at fp32-class accuracy: each operand is split into three bf16 pieces and the
 * six piece products that matter are accumulated in fp32 as one GEMM over K' = 6 K (csrc/gemm_f32.cu).  Used for the large
 * projections of the fp32 (decode) path.  ws: hnb_gemm_f32_tc_ws_bytes(M, N, K) bytes of scratch, 256-byte aligned. */
long long hnb_gemm_f32_tc_ws_bytes(int M, int N, int K);
int hnb_gemm_f32_tc(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                    int M, int N, int K, const float* bias, const float* R, long long ldr,
                    float* C, long long ldc, void* ws, void* stream);
/* on-device self test of the tcgen05 descriptor variants; returns 0 and fills max_abs_err[4] (host) */
int hnb_umma_selftest(float* max_abs_err_host, void* stream);

/* ---- ConvSubsampling4 front end (the step before the hot path) ---------------- */

/* First Conv2d(1 -> C, kernel 3, stride 2) + ReLU of ConvSubsampling4 (src/dcasr/models/encoder.py:55-70),
 * input [B, T, F] fp32 (one channel), weight [C, 9] fp32, bias [C] fp32.
 * out: bf16, logical [B, C, T1, F1] stored NHWC, i.e. [B, T1, F1, C] contiguous (T1 = (T-3)/2+1, F1 = (F-3)/2+1):
 * the layout the tensor-core kernels of the second convolution consume.  C % 8 == 0, C <= 1024. */
int hnb_subsample_conv1_fwd(const float* feats, const float* w, const float* bias, int B, int T, int F, int C,
                            void* out, void* stream);
/* The same with the output dtype chosen by the caller (HNB_BF16 as above, or HNB_F32: the fp32 no_grad forward of decoding,
 * src/dcasr/tasks/decode_task.py:123-151, where the reference runs this layer through cuDNN in fp32). */
int hnb_subsample_conv1_fwd_dt(const float* feats, const float* w, const float* bias, int B, int T, int F, int C,
                               void* out, int out_dtype, void* stream);
/* Backward of the same: a1 (the forward's output) and dout (both bf16, NHWC) -> dw [C, 9] and db [C], ACCUMULATED
 * into pre-zeroed, 16-byte aligned fp32 buffers.  ReLU mask = a1 > 0; the input needs no gradient. */
int hnb_subsample_conv1_bwd(const float* feats, const void* a1, const void* dout, int B, int T, int F, int C,
                            float* dw, float* db, void* stream);

/* bias + ReLU of the second convolution's output (encoder.py:61-62), bf16 NHWC viewed as [rows, C], IN PLACE */
int hnb_bias_relu_fwd(void* x, const float* bias, long long rows, int C, void* stream);
/* the same on an fp32 tensor (the fp32 no_grad forward of decoding); C % 4 == 0, x and bias 16-byte aligned */
int hnb_bias_relu_fwd_f32(float* x, const float* bias, long long rows, int C, void* stream);
/* its backward in one pass: dpre = dout * [out > 0] (bf16), db [C] ACCUMULATED (pre-zeroed fp32, 16-byte aligned) */
int hnb_bias_relu_bwd(const void* dout, const void* out, void* dpre, float* db, long long rows, int C, void* stream);

/* ---- CTC head (the step after the hot path) ------------------------------------ */

/* Reference: CTCHead.log_probs / CTCHead.loss / CTCHead.frame_argmax, src/dcasr/decoders/ctc.py:100-120:
 *   lp = log_softmax(Linear(features).float(), -1);  F.ctc_loss(lp^T, targets, feat_lengths, target_lengths, blank,
 *   reduction, zero_infinity=True).  The kernels work on the LOGITS as the projection wrote them (fp32 or bf16,
 *   [B, T, V1] with row stride ldl) and never materialise the fp32 log-probabilities.
 * hnb_ctc_lse: lse[row] = logsumexp over the V1 classes (fp32); argmax (optional, int32): per-frame top class, ties to the
 *   lowest id as torch.argmax (greedy decoding / frame_argmax). */
int hnb_ctc_lse(const void* logits, int dtype, long long rows, int V1, long long ldl, float* lse, int* argmax, void* stream);
/* alpha, beta [B, T, 2U+1] fp32 (log domain; both include the emission at t, PyTorch's convention) and nll [B] =
 * -log p(targets_b | logits_b) (inf when no alignment exists).  targets [B, U] int64 (padding beyond tgt_lens ignored),
 * feat_lens, tgt_lens [B] int64 (clamped to [0, T] and [0, U]).  U <= 511. */
int hnb_ctc_alpha_beta(const void* logits, int dtype, const float* lse, const long long* targets, const long long* feat_lens,
                       const long long* tgt_lens, int B, int T, int V1, long long ldl, int U, int blank, float* alpha,
                       float* beta, float* nll, void* stream);
/* dlogits [B, T, V1] (logits' dtype, row stride ldd) = gscale[b] * d nll[b] / d logits; zero beyond feat_lens[b] and for
 * utterances with nll = inf (zero_infinity).  gscale carries the upstream gradient and the reduction's weights. */
int hnb_ctc_grad(const void* logits, int dtype, const float* lse, const float* alpha, const float* beta,
                 const long long* targets, const long long* feat_lens, const long long* tgt_lens, const float* nll,
                 const float* gscale, int B, int T, int V1, long long ldl, int U, int blank, void* dlogits, long long ldd,
                 void* stream);

/* float64 twins of hnb_ema_fwd / _bwd (P float64 too) and hnb_window_reduce / _broadcast, for the reference's own
 * double-precision gradcheck tests of DynamicChunker._ema (tests/test_hnet_chunk.py:242-263) and of the fixed-stride pool
 * (tests/test_fixed_pool.py:197-207); one thread per (row, channel), test-sized tensors.  dP pre-zeroed. */
int hnb_ema_fwd_f64(const double* x, const double* P, int B, int M, int D, double p_clamp, double* out, void* stream);
int hnb_ema_bwd_f64(const double* dout, const double* x, const double* out, const double* P, int B, int M, int D,
                    double p_clamp, double* dx, double* dP, void* stream);
int hnb_window_reduce_f64(const double* x, const uint8_t* mask, int B, int L, int D, int M, int stride, int normalize,
                          double* z, float* cnt, void* stream);
int hnb_window_broadcast_f64(const double* z, const uint8_t* mask, const float* cnt, const double* resid, int B, int L,
                             int D, int M, int stride, double* out, void* stream);

/* out[c] += sum over rows of x[r, c] (fp32, ACCUMULATED into a pre-zeroed buffer): the bias gradient of nn.Linear
 * (CTCHead.proj, DCASREncoder.proj_in / proj_out: src/dcasr/models/encoder.py:104-111), x row-major with row stride ldx. */
int hnb_col_sum(const void* x, int dtype, long long rows, int cols, long long ldx, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HNET_B200_H */
