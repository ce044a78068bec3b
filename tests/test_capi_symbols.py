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
