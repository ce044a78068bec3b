"""Per-kernel timings at the headline shapes (B=40 x 16 s, Type A Small).  CUDA events, L2 flushed between reps."""
import sys, json
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from dcasr_b200 import ops, _lib
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

def timeit(fn, reps=3, inner=12):
    """`inner` back-to-back launches between one event pair: the host's per-call cost (tensor-map encode, ctypes,
    allocator) is hidden behind the queued GPU work, as it is inside a real step.  L2 is flushed before each rep."""
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort(); return ts[len(ts) // 2]

def bf(*shape): return (torch.randn(*shape, device=DEV) * 0.5).to(torch.bfloat16)
rows = []
def rec(name, us, flops=None, bytes_=None):
    r = {"kernel": name, "us": round(us, 1)}
    if flops: r["TFLOPs"] = round(flops / us / 1e6, 1)
    if bytes_: r["GBs"] = round(bytes_ / us / 1e3, 1)
    rows.append(r); print(r, flush=True)

for (T, d, di, H, tag) in ((15920, 384, 768, 12, "outer"), (7840, 512, 1024, 16, "main")):
    N = 128; C = di + 2 * N; dip = 2 * di + 2 * N + H; ds = (dip + 7) // 8 * 8; ldz = 2 * ds
    B = 40; L = T // B
    h = bf(T, d); Win = bf(ldz, d); zx = bf(T, ldz); yn = bf(T, 2 * di); Wout = bf(d, 2 * di); x = bf(T, d)
    rec(f"{tag} gemm in_proj fwd [{T}x{ldz}x{d}]", timeit(lambda: ops.gemm(h, Win)), 2 * T * ldz * d, T * (d + ldz) * 2)
    rec(f"{tag} gemm out_proj fwd+res [{T}x{d}x{2*di}]", timeit(lambda: ops.gemm(yn, Wout, residual=x)), 2 * T * d * 2 * di, T * (2 * di + 2 * d) * 2)
    rec(f"{tag} gemm in_proj dgrad [{T}x{d}x{ldz}]", timeit(lambda: ops.gemm(zx, Win, trans_b=True)), 2 * T * ldz * d, T * (d + ldz) * 2)
    rec(f"{tag} gemm out_proj dgrad [{T}x{2*di}x{d}]", timeit(lambda: ops.gemm(x, Wout, trans_b=True)), 2 * T * d * 2 * di, T * (2 * di + d) * 2)
    sk = ops.wgrad_splitk(T, ldz, d)
    rec(f"{tag} gemm in_proj wgrad sk{sk} [{ldz}x{d}x{T}]", timeit(lambda: ops.gemm(zx, h, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32)), 2 * T * ldz * d, T * (d + ldz) * 2)
    sk = ops.wgrad_splitk(T, d, 2 * di)
    rec(f"{tag} gemm out_proj wgrad sk{sk} [{d}x{2*di}x{T}]", timeit(lambda: ops.gemm(x, yn, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32)), 2 * T * d * 2 * di, T * (2 * di + d) * 2)
    g, b_ = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
    rec(f"{tag} layernorm_fwd", timeit(lambda: ops.layernorm_fwd(x, g, b_, 1e-5, torch.bfloat16)), None, T * d * 4)
    y_, mean, rstd = ops.layernorm_fwd(x, g, b_, 1e-5, torch.bfloat16)
    rec(f"{tag} layernorm_bwd", timeit(lambda: ops.layernorm_bwd(h, x, g, mean, rstd, x)), None, T * d * 8)
    cw, cb = torch.randn(2, C, 4, device=DEV), torch.randn(2, C, device=DEV)
    dtb, Al, Dk, nw = torch.randn(2, H, device=DEV), torch.log(torch.rand(2, H, device=DEV) * 15 + 1), torch.randn(2, H, device=DEV), torch.randn(2, di, device=DEV)
    lens = torch.full((B,), L, dtype=torch.int32, device=DEV)
    rec(f"{tag} conv_fwd", timeit(lambda: ops.conv_fwd(zx, ds, lens, cw, cb, dtb, 2, B, L, di, N, H)), None, 2 * T * C * 4)
    xconv, dt = ops.conv_fwd(zx, ds, lens, cw, cb, dtb, 2, B, L, di, N, H)
    rec(f"{tag} ssd_fwd tc", timeit(lambda: ops.ssd_fwd(xconv, dt, Al, Dk, 2, B, L, di, N, H, impl=1)), 4.0 * di * N * 2 * T, 2 * T * (C + di) * 2)
    yy, ws = ops.ssd_fwd(xconv, dt, Al, Dk, 2, B, L, di, N, H, impl=1)
    rec(f"{tag} gated_norm_fwd", timeit(lambda: ops.gated_norm_fwd(yy, zx, ds, lens, nw, 2, B, L, di)), None, 2 * T * di * 6)
    yn2, rs = ops.gated_norm_fwd(yy, zx, ds, lens, nw, 2, B, L, di)
    dzx = torch.zeros_like(zx)
    rec(f"{tag} gated_norm_bwd", timeit(lambda: ops.gated_norm_bwd(yn, yy, zx, ds, lens, nw, rs, 2, B, L, di, dzx)), None, 2 * T * di * 10)
    dy = bf(2, T, di)
    _lib.profile_start(); ops.ssd_bwd(dy, xconv, yy, dt, Al, Dk, ws, 2, B, L, di, N, H, impl=1, keep_parts=True); _lib.profile_stop()
    rec(f"{tag} ssd_bwd tc (3 kernels)", timeit(lambda: ops.ssd_bwd(dy, xconv, yy, dt, Al, Dk, ws, 2, B, L, di, N, H, impl=1, keep_parts=True)), 2.5 * 4.0 * di * N * 2 * T, 2 * T * (C * 2 + di * 3) * 2)
    dxc, dBC, ddt, dA, dD = ops.ssd_bwd(dy, xconv, yy, dt, Al, Dk, ws, 2, B, L, di, N, H, impl=1, keep_parts=True)
    rec(f"{tag} conv_bwd", timeit(lambda: ops.conv_bwd(zx, dxc, dBC, ddt, ds, lens, cw, cb, dtb, 2, B, L, di, N, H, dzx)), None, 2 * T * C * 6)
json.dump(rows, open("gpurun_out/kbench.json", "w"), indent=1)
