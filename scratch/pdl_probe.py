"""Does the programmatic-launch attribute do anything?  A chain of short layernorm launches (a few us each) timed with
CUDA events on torch's default stream and on a side stream; run with HNB_PDL=1 and HNB_PDL=0."""
import os, sys, ctypes
sys.path.insert(0, "tests"); sys.path.insert(0, "h-net-mamba-asr_b200")
import torch
from dcasr_b200 import ops, _lib
L_ = _lib.lib()
DEV = "cuda"
for rows in (512, 15920):
    x = torch.randn(rows, 384, device=DEV).bfloat16(); g = torch.randn(384, device=DEV); b = torch.randn(384, device=DEV)
    y = torch.empty_like(x); mean = torch.empty(rows, device=DEV); rstd = torch.empty(rows, device=DEV)
    def chain(n, stream):
        f = L_.raw("layernorm_fwd")
        for _ in range(n):
            f(x.data_ptr(), 1, g.data_ptr(), b.data_ptr(), rows, 384, ctypes.c_float(1e-5), y.data_ptr(), 1, mean.data_ptr(), rstd.data_ptr(), stream)
    for name, s in (("default", torch.cuda.current_stream()), ("side", torch.cuda.Stream())):
        with torch.cuda.stream(s):
            chain(50, s.cuda_stream); torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s); chain(2000, s.cuda_stream); e1.record(s); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 2000 * 1e3)
            print(f"HNB_PDL={os.environ.get('HNB_PDL','1')} rows={rows} stream={name} ({s.cuda_stream}): {min(ts):.2f} us per launch", flush=True)
