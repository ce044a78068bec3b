"""ctypes binding of libhnet_b200.so, generated from include/hnet_b200.h.

The product path has NO fallback: if the shared library is missing, or a kernel call is made with a
non-CUDA tensor, this module raises.  Signatures are parsed from the header so the binding cannot
drift from the declared C ABI.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhnet_b200.so")
HEADER = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "hnet_b200.h")

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}


class HnbError(RuntimeError):
    pass


def dtype_code(t: torch.dtype) -> int:
    try:
        return _DT[t]
    except KeyError:
        raise HnbError(f"hnet_b200 kernels take float32 or bfloat16 activations, got {t}") from None


def parse_header(path: str = HEADER) -> dict[str, tuple[str, list[str]]]:
    """name -> (return type, [arg C types]) for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?[\w ]+?\**)\s*(hnb_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if args in ("void", ""):
            arg_types = []
        else:
            arg_types = []
            for a in args.split(","):
                a = a.strip()
                ty = re.sub(r"\s*\w+$", "", a) if not a.endswith("*") else a   # drop the parameter name
                arg_types.append(ty.strip())
        out[name] = (ret, arg_types)
    return out


def _ctype(ty: str):
    ty = ty.replace("const ", "").strip()
    if ty.endswith("*"):
        return ctypes.c_char_p if ty == "char*" else ctypes.c_void_p
    return {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
            "void": None}[ty]


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise HnbError(
                f"{LIB_PATH} is missing: build it with `python h-net-mamba-asr_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.decls = parse_header()
        self.fn = {}
        for name, (ret, args) in self.decls.items():
            f = getattr(self.cdll, name)
            f.restype = _ctype(ret)
            f.argtypes = [_ctype(a) for a in args]
            self.fn[name] = f

    def raw(self, name: str):
        return self.fn["hnb_" + name]

    def call(self, name: str, *args) -> None:
        """Call int hnb_<name>(...) with tensors turned into device pointers; raise on non-zero status.  The stream
        argument (``stream()``) resolves to torch's current stream on the device of the call's first tensor, and the
        call runs with that device current: a model on "cuda:1" in a process whose current device is 0 works."""
        conv = []
        T = torch.Tensor
        dev, spos = -1, -1
        for a in args:
            if type(a) is T or isinstance(a, T):                  # (exact-type test first: it is the common case)
                if not a.is_cuda:
                    raise HnbError(f"hnb_{name}: got a {a.device} tensor; the hot path is CUDA-only")
                if not a.is_contiguous():
                    raise HnbError(f"hnb_{name}: tensor argument must be contiguous")
                if dev < 0:
                    dev = a.device.index
                conv.append(a.data_ptr())
            elif a is _CURRENT:
                spos = len(conv)
                conv.append(0)
            else:
                conv.append(a)
        cur = torch.cuda.current_device()
        if dev >= 0 and dev != cur:                               # rare: tensors on another device than the current one
            with torch.cuda.device(dev):
                return self._run(name, conv, spos, dev, args)
        return self._run(name, conv, spos, cur, args)

    def _run(self, name, conv, spos, dev, args) -> None:
        if spos >= 0:
            conv[spos] = _raw_stream(dev) if _raw_stream is not None else torch.cuda.current_stream(dev).cuda_stream
        prof = _PROFILE
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self.fn["hnb_" + name](*conv)
            e1.record()
            prof.append((name, tuple(a for a in args if isinstance(a, (int, float))), e0, e1))
        else:
            rc = self.fn["hnb_" + name](*conv)
        if rc != 0:
            raise HnbError(f"hnb_{name} failed ({rc}): {self.cdll.hnb_last_error().decode()}")


_LIB: _Lib | None = None
_PROFILE: list | None = None      # when a list: every call appends (name, scalar args, start event, end event)


def profile_start() -> None:
    global _PROFILE
    _PROFILE = []


def profile_stop() -> list:
    """-> [(kernel entry point, scalar args, milliseconds)] for every call since profile_start()."""
    global _PROFILE
    rec, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in rec]


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
        _LIB.cdll.hnb_last_error.restype = ctypes.c_char_p
    return _LIB


class _CurrentStream:
    """Placeholder for "torch's current stream on the device of this call's tensors"; resolved inside _Lib.call.
    Handed straight to a ctypes function (``lib().raw(...)``) it converts to the current device's current stream."""
    __slots__ = ()

    @property
    def _as_parameter_(self):
        dev = torch.cuda.current_device()
        return ctypes.c_void_p(_raw_stream(dev) if _raw_stream is not None else torch.cuda.current_stream(dev).cuda_stream)


_CURRENT = _CurrentStream()
# One C call; building a torch.cuda.Stream object per kernel launch (device-index lookups, availability checks) cost
# 15 us of host time per launch, 5 ms of the 15 ms the host needed to enqueue one encoder step.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> _CurrentStream:
    return _CURRENT


def launch_count() -> int:
    return int(lib().raw("launch_count")())


def reset_launch_count() -> None:
    lib().raw("reset_launch_count")()
