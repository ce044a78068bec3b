"""Low-level wrappers over the C ABI (include/hnet_b200.h): tensor plumbing only, no arithmetic.

Every function allocates its outputs with torch (PyTorch owns device memory), enqueues the kernels on
torch's current CUDA stream through ctypes, and returns tensors.  There is no CPU path: a CPU tensor
raises HnbError.
"""
from __future__ import annotations

import torch

from ._lib import BF16, F32, HnbError, dtype_code, lib, stream

# ssd implementation selector: 0 = CUDA-core fp32 (exact), 1 = tcgen05 (bf16). "auto" picks by dtype.
SSD_IMPL = "auto"
# dense projections: "tcgen05" (own kernel) for bf16; fp32 always uses the exact CUDA-core GEMM
GEMM_BF16_IMPL = "tcgen05"
# fp32 projections: "tc" = large ones as bf16-piece GEMMs on the tensor cores (fp32-class accuracy), "exact" = CUDA cores only
GEMM_F32_IMPL = "tc"


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------
# dense projections
# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, bias=None, residual=None,
         out_dtype=None, splitk=1, out=None) -> torch.Tensor:
    """C[M,N] = op(a) @ op(b)^T-convention of hnb_gemm_*:
       trans_a=False: a is [M,K];  True: a is [K,M].   trans_b=False: b is [N,K] (nn.Linear weight);  True: b is [K,N].
       2-D tensors whose rows are contiguous (stride(1) == 1); row strides are honoured."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != Kb:
        raise HnbError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    L = lib()
    if not (a.is_cuda and b.is_cuda):
        raise HnbError("gemm: the hot path is CUDA-only")

    def ptr(t):                        # row-strided views (column blocks of a wider matrix) go in by address + row stride
        return t if t is None or t.is_contiguous() else t.data_ptr()
    if a.dtype == torch.float32:
        if b.dtype != torch.float32:
            raise HnbError("gemm: mixed operand dtypes")
        c = out if out is not None else _empty((M, N), torch.float32, a)
        r = residual
        if GEMM_F32_IMPL == "tc" and float(M) * N * K >= 67108864.0:
            # large fp32 projections (decode): bf16-piece GEMM on the tensor cores, fp32-class accuracy (csrc/gemm_f32.cu)
            ws = torch.empty(int(L.raw("gemm_f32_tc_ws_bytes")(M, N, K)), dtype=torch.uint8, device=a.device)
            L.call("gemm_f32_tc", ptr(a), a.stride(0), int(trans_a), ptr(b), b.stride(0), int(trans_b), M, N, K,
                   bias, ptr(r), r.stride(0) if r is not None else 0, ptr(c), c.stride(0), ws, stream())
            return c
        L.call("gemm_f32", ptr(a), a.stride(0), int(trans_a), ptr(b), b.stride(0), int(trans_b), M, N, K,
               bias, ptr(r), r.stride(0) if r is not None else 0, ptr(c), c.stride(0), 0, stream())
        return c
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise HnbError(f"gemm: unsupported operand dtypes {a.dtype}, {b.dtype}")
    od = out_dtype or torch.bfloat16
    if splitk > 1:
        c = out if out is not None else torch.zeros((M, N), dtype=torch.float32, device=a.device)
        od = torch.float32
    else:
        c = out if out is not None else _empty((M, N), od, a)
    r = residual
    if r is not None and r.dtype != od:
        raise HnbError("gemm: residual dtype must equal the output dtype")
    L.call("gemm_bf16", ptr(a), a.stride(0), int(trans_a), ptr(b), b.stride(0), int(trans_b), M, N, K, bias,
           ptr(r), r.stride(0) if r is not None else 0, ptr(c), c.stride(0), dtype_code(od), int(splitk), stream())
    return c


def col_sum(x: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a 2-D activation (rows of stride x.stride(0)): the bias gradient of nn.Linear."""
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    lib().call("col_sum", x.data_ptr(), dtype_code(x.dtype), x.shape[0], x.shape[1], x.stride(0), out, stream())
    return out


def wgrad_splitk(k_tokens: int, m: int, n: int) -> int:
    """split-K factor for weight gradients (K = tokens), chosen by the library for the tile grid it will use."""
    return max(1, int(lib().raw("gemm_splitk_hint")(int(m), int(n), int(k_tokens))))


# ------------------------------------------------------------------------------------------------
# H-Net stage
# ------------------------------------------------------------------------------------------------
def router_fwd(qk, mask_u8, B, L, D, pb_dtype, N):
    L_ = lib()
    n = B * L
    p = _empty((B, L), pb_dtype, qk)
    b = _empty((B, L), pb_dtype, qk)
    nblk = L_.raw("router_num_partials")(n)
    partial = _empty((nblk * 4,), torch.float32, qk)
    stats = _empty((8,), torch.float32, qk)
    L_.call("router_fwd", qk, dtype_code(qk.dtype), qk.stride(0), mask_u8, B, L, D, 1e-6, p, b,
            dtype_code(pb_dtype), partial, stream())
    L_.call("ratio_finalize", partial, nblk, float(N), stats, stream())
    return p, b, stats


def router_bwd(qk, mask_u8, B, L, D, dp, dratio, stats, N):
    dqk = torch.empty_like(qk)
    lib().call("router_bwd", qk, dtype_code(qk.dtype), qk.stride(0), mask_u8, B, L, D, 1e-6, dp, dratio, stats,
               float(N), dqk, stream())
    return dqk


def boundary_scan(b, B, L):
    L_ = lib()
    memb = _empty((B, L), torch.int64, b)
    counts = _empty((B,), torch.int32, b)
    ws = _empty((L_.raw("boundary_scan_ws_bytes")(B * L),), torch.uint8, b)
    L_.call("boundary_scan", b, dtype_code(b.dtype), B, L, memb, counts, ws, stream())
    return memb, counts


def compact_rows(x, p, b, memb, counts, M):
    B, L, D = x.shape
    z = _empty((B, M, D), x.dtype, x)
    zmask = _empty((B, M), torch.uint8, x)
    P = _empty((B, M), torch.float32, x)
    starts = _empty((B, M), torch.int32, x)
    lib().call("compact_rows", x, dtype_code(x.dtype), p, b, dtype_code(p.dtype), memb, counts, B, L, D, M,
               z, zmask, P, starts, stream())
    return z, zmask, P, starts


def compact_rows_bwd(dz, b, memb, L, dx=None):
    B, M, D = dz.shape
    acc = dx is not None
    if dx is None:
        dx = _empty((B, L, D), dz.dtype, dz)
    lib().call("compact_rows_bwd", dz, dtype_code(dz.dtype), b, dtype_code(b.dtype), memb, B, L, D, M, dx,
               int(acc), stream())
    return dx


def ema_fwd(x, P, p_clamp=1e-4):
    B, M, D = x.shape
    out = torch.empty_like(x)
    if x.dtype == torch.float64:                    # the reference's double-precision gradchecks (csrc/f64_kernels.cu)
        lib().call("ema_fwd_f64", x, P, B, M, D, float(p_clamp), out, stream())
        return out
    lib().call("ema_fwd", x, dtype_code(x.dtype), P, B, M, D, float(p_clamp), out, stream())
    return out


def ema_bwd(dout, x, out, P, p_clamp=1e-4):
    B, M, D = x.shape
    dx = torch.empty_like(x)
    dP = torch.zeros_like(P)
    if x.dtype == torch.float64:
        lib().call("ema_bwd_f64", dout, x, out, P, B, M, D, float(p_clamp), dx, dP, stream())
        return dx, dP
    lib().call("ema_bwd", dout, x, out, dtype_code(x.dtype), P, B, M, D, float(p_clamp), dx, dP, stream())
    return dx, dP


def upsample_fwd(zbar, memb, p, b, resid, y_dtype):
    B, M, D = zbar.shape
    L = memb.shape[1]
    y = _empty((B, L, D), y_dtype, zbar)
    lib().call("upsample_fwd", zbar, dtype_code(zbar.dtype), memb, p, b, dtype_code(p.dtype), resid, B, L, D, M,
               y, dtype_code(y_dtype), stream())
    return y


def upsample_bwd(dy, zbar, memb, starts, counts, p, b):
    B, M, D = zbar.shape
    L = memb.shape[1]
    dz = torch.empty_like(zbar)
    dp = _empty((B, L), torch.float32, zbar)
    lib().call("upsample_bwd", dy, dtype_code(dy.dtype), zbar, dtype_code(zbar.dtype), memb, starts, counts, p, b,
               dtype_code(p.dtype), B, L, D, M, dz, dp, stream())
    return dz, dp


def masked_ratio_stats(p, b, mask_u8, N):
    L_ = lib()
    n = p.numel()
    nblk = (n + 255) // 256
    partial = _empty((nblk * 4,), torch.float32, p)
    stats = _empty((8,), torch.float32, p)
    L_.call("masked_sums", p, b, dtype_code(p.dtype), mask_u8, n, partial, stream())
    L_.call("ratio_finalize", partial, nblk, float(N), stats, stream())
    return stats


# ------------------------------------------------------------------------------------------------
# Mamba-2 block
# ------------------------------------------------------------------------------------------------
def layernorm_fwd(x2, gamma, beta, eps, y_dtype):
    rows, d = x2.shape
    y = _empty((rows, d), y_dtype, x2)
    mean = _empty((rows,), torch.float32, x2)
    rstd = _empty((rows,), torch.float32, x2)
    lib().call("layernorm_fwd", x2, dtype_code(x2.dtype), gamma, beta, rows, d, float(eps), y, dtype_code(y_dtype),
               mean, rstd, stream())
    return y, mean, rstd


def layernorm_bwd(dy, x2, gamma, mean, rstd, dres, acc=None):
    """acc: optional pre-zeroed fp32 [2, d] accumulator for (dgamma, dbeta)."""
    rows, d = x2.shape
    dx = torch.empty_like(x2)
    if acc is None:
        acc = torch.zeros(2, d, dtype=torch.float32, device=x2.device)
    dg, db = acc[0], acc[1]
    lib().call("layernorm_bwd", dy, dtype_code(dy.dtype), x2, dtype_code(x2.dtype), gamma, mean, rstd, dres, rows, d,
               dx, dtype_code(dx.dtype), dg, db, stream())
    return dx, dg, db


def conv_fwd(zx, dstride, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H):
    C = di + 2 * N
    xconv = _empty((ndir, B * L, C), zx.dtype, zx)
    dt = _empty((ndir, B * L, H), torch.float32, zx)
    lib().call("conv_fwd", zx, dtype_code(zx.dtype), zx.stride(0), dstride, lengths, conv_w, conv_b, dt_bias,
               ndir, B, L, di, N, H, xconv, dt, stream())
    return xconv, dt


def conv_bwd(zx, dxc, dBC, ddt, dstride, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H, dzx, acc=None):
    """acc: optional pre-zeroed (dconv_w, dconv_b, ddt_bias) accumulators."""
    if acc is None:
        acc = (torch.zeros_like(conv_w), torch.zeros_like(conv_b), torch.zeros_like(dt_bias))
    dw, db, ddtb = acc
    parts = dBC.shape[0] if dBC.dim() == 4 else 1          # [parts, ndir, B*L, 2N]: partial sums of the SSD dB/dC kernel
    lib().call("conv_bwd", zx, dxc, dtype_code(zx.dtype), zx.stride(0), dstride, dBC, ddt, lengths, conv_w, conv_b,
               dt_bias, ndir, B, L, di, N, H, dzx, dw, db, ddtb, parts, stream())
    return dw, db, ddtb


def ssd_impl_for(dtype) -> int:
    if SSD_IMPL == "auto":
        return 1 if dtype == torch.bfloat16 else 0      # tcgen05 for bf16, exact CUDA-core path for fp32
    return int(SSD_IMPL)


def ssd_fwd(xconv, dt, A_log, D, ndir, B, L, di, N, H, impl=None):
    L_ = lib()
    impl = ssd_impl_for(xconv.dtype) if impl is None else impl
    y = _empty((ndir, B * L, di), xconv.dtype, xconv)
    ws = _empty((L_.raw("ssd_ws_bytes")(ndir, B, L, di, N, H) // 4,), torch.float32, xconv)
    L_.call("ssd_fwd", xconv, dtype_code(xconv.dtype), dt, A_log, D, ndir, B, L, di, N, H, y, ws, impl, stream())
    return y, ws


def ssd_bwd(dy, xconv, y, dt, A_log, D, ws, ndir, B, L, di, N, H, impl=None, acc=None, keep_parts=False):
    """acc: optional pre-zeroed (dA_log, dD) accumulators.  keep_parts: return dBC as [parts, ndir, B*L, 2N] partial sums
    (what conv_bwd consumes) instead of their sum."""
    L_ = lib()
    impl = ssd_impl_for(dy.dtype) if impl is None else impl
    dxc = torch.empty_like(dy)
    parts = int(L_.raw("ssd_dbc_parts")(ndir, B, L, H, impl))
    dBC = _empty((parts, ndir, B * L, 2 * N), dy.dtype, dy)
    ddt = torch.empty_like(dt)
    dA, dD = acc if acc is not None else (torch.zeros_like(A_log), torch.zeros_like(D))
    ws2 = _empty((L_.raw("ssd_ws_bytes")(ndir, B, L, di, N, H) // 4,), torch.float32, dy)
    L_.call("ssd_bwd", dy, xconv, y, dtype_code(dy.dtype), dt, A_log, D, ws, ndir, B, L, di, N, H, dxc, dBC, parts, ddt,
            dA, dD, ws2, impl, stream())
    if not keep_parts:
        dBC = dBC[0] if parts == 1 else dBC.float().sum(0).to(dy.dtype)
    return dxc, dBC, ddt, dA, dD


def gated_norm_fwd(y, zx, dstride, lengths, norm_w, ndir, B, L, di, eps=1e-5):
    out = _empty((B * L, ndir * di), y.dtype, y)
    rstd = _empty((ndir, B * L), torch.float32, y)
    lib().call("gated_norm_fwd", y, zx, dtype_code(y.dtype), zx.stride(0), dstride, lengths, norm_w, ndir, B, L, di,
               float(eps), out, rstd, stream())
    return out, rstd


def gated_norm_bwd(dout, y, zx, dstride, lengths, norm_w, rstd, ndir, B, L, di, dzx, acc=None):
    dy = torch.empty_like(y)
    dw = acc if acc is not None else torch.zeros_like(norm_w)
    lib().call("gated_norm_bwd", dout, y, zx, dtype_code(y.dtype), zx.stride(0), dstride, lengths, norm_w, rstd,
               ndir, B, L, di, dy, dzx, dw, stream())
    return dy, dw


def pack_mixer_params(params, ndir, d, di, N, H, dstride, w_dtype, device):
    """params: per direction (in_proj.w, conv1d.w, conv1d.b, dt_bias, A_log, D, norm.w, out_proj.w), fp32 CUDA.
    -> Win [ndir*dstride, d], Wout [d, ndir*di] (w_dtype) and the fp32 stacks, with ndir kernel launches."""
    C = di + 2 * N
    Win = torch.empty((ndir * dstride, d), dtype=w_dtype, device=device)
    Wout = torch.empty((d, ndir * di), dtype=w_dtype, device=device)
    small = torch.empty(ndir * (C * 5 + 3 * H + di), dtype=torch.float32, device=device)
    conv_w, conv_b, dt_bias, A_log, Dk, norm_w = small.split(
        [ndir * C * 4, ndir * C, ndir * H, ndir * H, ndir * H, ndir * di])
    L_ = lib()
    for t in params:
        if t.dtype != torch.float32:
            raise HnbError("mixer parameters must be fp32 masters")
    if ndir == 2:                                                  # both directions in one launch
        a, b = params[:8], params[8:16]
        L_.call("pack_mixer_params2", a[0], a[7], a[1], a[2], a[3], a[4], a[5], a[6],
                b[0], b[7], b[1], b[2], b[3], b[4], b[5], b[6], d, di, N, H, dstride, Win, Wout,
                dtype_code(w_dtype), conv_w, conv_b, dt_bias, A_log, Dk, norm_w, stream())
    else:
        for r in range(ndir):
            inw, cw, cb, dtb, al, dk, nw, outw = params[r * 8:(r + 1) * 8]
            L_.call("pack_mixer_params", inw, outw, cw, cb, dtb, al, dk, nw, r, ndir, d, di, N, H, dstride, Win, Wout,
                    dtype_code(w_dtype), conv_w, conv_b, dt_bias, A_log, Dk, norm_w, stream())
    return (Win, Wout, conv_w.view(ndir, C, 4), conv_b.view(ndir, C), dt_bias.view(ndir, H), A_log.view(ndir, H),
            Dk.view(ndir, H), norm_w.view(ndir, di))


# ------------------------------------------------------------------------------------------------
# ConvSubsampling4 front end
# ------------------------------------------------------------------------------------------------
def subsample_conv1_supported(C: int) -> bool:
    return C % 8 == 0 and C <= 1024


def subsample_conv1_fwd(feats, w, b, out_dtype=torch.bfloat16):
    """feats [B, T, F] fp32, w [C, 1, 3, 3] fp32, b [C] -> relu(conv) as bf16 (training) or fp32 (decode) [B, C, T1, F1] in
    channels_last memory."""
    B, T, F = feats.shape
    C = w.shape[0]
    T1, F1 = (T - 3) // 2 + 1, (F - 3) // 2 + 1
    out = torch.empty((B, T1, F1, C), dtype=out_dtype, device=feats.device)            # NHWC storage
    if out_dtype == torch.bfloat16:
        lib().call("subsample_conv1_fwd", feats, w, b, B, T, F, C, out, stream())
    else:
        lib().call("subsample_conv1_fwd_dt", feats, w, b, B, T, F, C, out, dtype_code(out_dtype), stream())
    return out.permute(0, 3, 1, 2)                                                      # logical NCHW, channels_last strides


def subsample_conv1_bwd(feats, a1, dout):
    """a1 (forward output), dout: bf16 [B, C, T1, F1] channels_last -> (dw [C, 1, 3, 3], db [C]) fp32."""
    B, T, F = feats.shape
    C = a1.shape[1]
    acc = torch.zeros(C * 10, dtype=torch.float32, device=feats.device)
    dw, db = acc[:C * 9], acc[C * 9:]
    lib().call("subsample_conv1_bwd", feats, a1.permute(0, 2, 3, 1), dout.permute(0, 2, 3, 1), B, T, F, C, dw, db, stream())
    return dw.view(C, 1, 3, 3), db


def bias_relu_fwd_(x2d, bias):
    """x2d [rows, C] bf16 (or fp32) contiguous (an NHWC tensor flattened), in place: relu(x + bias)."""
    rows, C = x2d.shape
    if x2d.dtype == torch.float32:                      # decode: the fp32 twin
        lib().call("bias_relu_fwd_f32", x2d, bias, rows, C, stream())
    else:
        lib().call("bias_relu_fwd", x2d, bias, rows, C, stream())
    return x2d


def bias_relu_bwd(dout2d, out2d):
    rows, C = out2d.shape
    dpre = torch.empty_like(out2d)
    db = torch.zeros(C, dtype=torch.float32, device=out2d.device)
    lib().call("bias_relu_bwd", dout2d, out2d, dpre, db, rows, C, stream())
    return dpre, db



# ------------------------------------------------------------------------------------------------
# FixedPoolChunker
# ------------------------------------------------------------------------------------------------
def window_reduce(x, mask_u8, M, stride, normalize, z_dtype, want_cnt=True):
    """x [B,L,D] -> z [B,M,D] (sum or masked mean over fixed windows), cnt [B,M] float (valid frames per window)."""
    B, L, D = x.shape
    z = _empty((B, M, D), z_dtype, x)
    cnt = _empty((B, M), torch.float32, x) if want_cnt else None
    if x.dtype == torch.float64 and z_dtype == torch.float64:
        lib().call("window_reduce_f64", x, mask_u8, B, L, D, M, int(stride), int(bool(normalize)), z, cnt, stream())
        return z, cnt
    lib().call("window_reduce", x, dtype_code(x.dtype), mask_u8, B, L, D, M, int(stride), int(bool(normalize)), z,
               dtype_code(z_dtype), cnt, stream())
    return z, cnt


def window_broadcast(z, mask_u8, cnt, resid, L, stride, out_dtype):
    """z [B,M,D] -> out [B,L,D]: out[b,t] = m[b,t] / max(cnt[b,w],1) * z[b,w(t)] (+ resid), w(t) = min(t//stride, M-1)."""
    B, M, D = z.shape
    out = _empty((B, L, D), out_dtype, z)
    if z.dtype == torch.float64 and out_dtype == torch.float64:
        lib().call("window_broadcast_f64", z, mask_u8, cnt, resid, B, L, D, M, int(stride), out, stream())
        return out
    lib().call("window_broadcast", z, dtype_code(z.dtype), mask_u8, cnt, resid, B, L, D, M, int(stride), out,
               dtype_code(out_dtype), stream())
    return out
