"""CPU: the C-ABI library builds, loads, and exports every symbol include/hnet_b200.h declares."""
import os
import subprocess

import pytest

from _util import PKG_DIR, REPO


def _lib_path():
    return os.path.join(PKG_DIR, "dcasr_b200", "libhnet_b200.so")


def test_library_builds_and_exports_declared_symbols():
    import importlib.util
    spec = importlib.util.spec_from_file_location("hnb_build", os.path.join(PKG_DIR, "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    b.build(verbose=False)
    from dcasr_b200._lib import lib, parse_header
    decls = parse_header()
    assert len(decls) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", _lib_path()], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = sorted(set(decls) - exported)
    assert not missing, f"declared in include/hnet_b200.h but not exported: {missing}"
    L = lib()                                     # ctypes load + argtypes for every declaration
    assert L.raw("version")() >= 100
    assert L.raw("ssd_chunk")() in (64, 128)
    assert L.raw("ssd_ws_bytes")(2, 40, 398, 768, 128, 12) > 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    obj = os.path.join(PKG_DIR, "build", "gemm_tcgen05.o")
    if not os.path.exists(obj):
        pytest.skip("objects not built")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, f"{mnem} missing: the GEMM is not on the tcgen05/TMA path"
    assert "HMMA.16816" not in sass, "legacy mma.sync path present"


def test_calls_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dcasr_b200 as d
    ch = d.DynamicChunker(16, N=2)
    with pytest.raises(Exception):
        ch.chunk(torch.randn(1, 8, 16))
    with pytest.raises(Exception):
        d.MambaBlock(64)(torch.randn(1, 8, 64))


def test_block_composite_layout_queries():
    """Size / layout queries of the one-call-per-block composites are pure host arithmetic (no GPU needed): workspaces
    are non-empty and grow with the batch, the gradient arena's offsets are increasing and 16-byte aligned."""
    import ctypes
    from dcasr_b200._lib import lib
    L_ = lib()
    B, L, d, ndir, di, N, H = 4, 398, 384, 2, 768, 128, 12
    f1 = L_.raw("block_fwd_ws_bytes")(B, L, d, ndir, di, N, H, 1)
    f2 = L_.raw("block_fwd_ws_bytes")(2 * B, L, d, ndir, di, N, H, 1)
    b1 = L_.raw("block_bwd_ws_bytes")(B, L, d, ndir, di, N, H, 1, 1)
    assert 0 < f1 < f2 and b1 > 0 and f1 % 256 == 0
    offs = (ctypes.c_longlong * 9)()
    n = L_.raw("block_grad_floats")(B, L, d, ndir, di, N, H, ctypes.addressof(offs))
    o = list(offs)
    dstride = (2 * di + 2 * N + H + 7) // 8 * 8
    assert o[0] == 0 and o[1] == d * ndir * di and o[2] == o[1] + ndir * dstride * d
    assert all(a < b for a, b in zip(o, o[1:])) and all(v % 4 == 0 for v in o[:3])
    assert n == o[8] + 2 * d


def test_host_side_cost_models_are_sane():
    """The library's launch-shaping heuristics are pure host arithmetic (they fall back to 148 SMs without a GPU):
    dB/dC head groups are 1 or 2 and divide H, the CUDA-core path always asks for one part, split-K hints are >= 1
    and never exceed the number of 64-wide k-blocks, and the SSD workspace grows with the problem."""
    from dcasr_b200._lib import lib
    L_ = lib()
    parts = L_.raw("ssd_dbc_parts")
    for (ndir, B, L, H) in ((2, 40, 398, 12), (2, 40, 196, 16), (2, 2, 1498, 24), (1, 1, 7, 3), (2, 3, 150, 5)):
        p1, p3 = parts(ndir, B, L, H, 1), parts(ndir, B, L, H, 3)
        assert p1 == 1, (ndir, B, L, H, p1)          # fused backward: cut items are summed inside the kernel (fix-up buffer)
        assert p3 in (1, 2) and H % p3 == 0, (ndir, B, L, H, p3)
        assert parts(ndir, B, L, H, 0) == 1
    assert parts(2, 40, 196, 16, 3) == 2            # three-kernel backward, main stack at the headline batch: 160 items are 1.1 waves of 148 SMs
    hint = L_.raw("gemm_splitk_hint")
    for (M, N, K) in ((3616, 384, 15920), (384, 1536, 15920), (4640, 512, 7840), (128, 128, 64), (8, 8, 8)):
        sk = hint(M, N, K)
        assert 1 <= sk <= max(1, (K + 63) // 64), (M, N, K, sk)
    assert hint(3616, 384, 15920) > 1               # weight gradients (K = tokens) are split
    ws = L_.raw("ssd_ws_bytes")
    assert 0 < ws(2, 4, 398, 768, 128, 12) < ws(2, 8, 398, 768, 128, 12) < ws(2, 8, 1498, 768, 128, 12)


def test_ssd_span_schedule_cuts():
    """The fused SSD backward cuts the sequence of all (row-chunk, head) steps into one piece per SM.  Host arithmetic only:
    the cuts are monotone, cover every step exactly once, never fall twice inside one item (so dB | dC has at most two
    partial sums per row-chunk), and the pieces carry equal cost within one head-step plus the snapping slack."""
    import ctypes
    from dcasr_b200._lib import lib
    f = lib().raw("ssd_span_cuts")
    for (ndirB, L, H, G) in ((80, 398, 12, 148), (80, 196, 16, 148), (20, 1498, 16, 148), (20, 640, 24, 148), (6, 398, 12, 148),
                             (1, 7, 3, 148), (4, 150, 5, 132), (80, 256, 12, 148), (3, 129, 4, 16)):
        out = (ctypes.c_int * (G + 1))()
        assert f(ndirB, L, H, G, out) == 0
        c = list(out)
        nc = (L + 127) // 128
        steps = ndirB * nc * H
        assert c[0] == 0 and c[-1] == steps and all(a <= b for a, b in zip(c, c[1:])), (ndirB, L, H, c[:8])
        inner = {}
        for v in c[1:-1]:
            if v % H and v < steps:
                inner[v // H] = inner.get(v // H, set()) | {v}
        assert all(len(v) == 1 for v in inner.values()), (ndirB, L, H)
        if steps >= 8 * G:                           # big problems: the pieces are balanced
            rem = L - (nc - 1) * 128
            nfull = ndirB * (nc if rem == 128 else nc - 1) * H
            w = 0.9 + 0.1 * rem / 128
            cost = lambda a, b: max(0, min(b, nfull) - a) + w * max(0, b - max(a, nfull))
            costs = [cost(a, b) for a, b in zip(c, c[1:])]
            ideal = sum(costs) / G
            assert max(costs) <= ideal + (1.5 if ideal >= H else H / 2 + 1.5), (ndirB, L, H, ideal, max(costs))
