"""ConvSubsampling4 front-end kernels alone at the headline shape (40 x 16 s, C = 384): CUDA events, L2 flushed."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, reps=5, inner=3):
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
torch.manual_seed(0)
B, T, F, C = 40, 1598, 80, 384
feats = torch.randn(B, T, F, device=DEV)
w = torch.randn(C, 1, 3, 3, device=DEV) * 0.3; b = torch.randn(C, device=DEV) * 0.1
out = ops.subsample_conv1_fwd(feats, w, b)
g = torch.randn_like(out)
us_f = timeit(lambda: ops.subsample_conv1_fwd(feats, w, b))
us_b = timeit(lambda: ops.subsample_conv1_bwd(feats, out, g))
dw, db = ops.subsample_conv1_bwd(feats, out, g)
nbytes = out.numel() * 2
print(f"env {os.environ.get('HNB_SUB_FWD_CPT', '-')}/{os.environ.get('HNB_SUB_BWD_THREADS', '-')}: fwd {us_f:.0f} us ({nbytes / us_f / 1e3:.0f} GB/s)  bwd {us_b:.0f} us ({2 * nbytes / us_b / 1e3:.0f} GB/s)  "
      f"checks {float(out.float().abs().mean()):.6f} {float(dw.abs().mean()):.4f} {float(db.abs().mean()):.4f}")
