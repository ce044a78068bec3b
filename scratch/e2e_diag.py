"""Where does the e2e leg lose time on some boxes?  H2D copy bandwidth from pinned memory, alone and next to compute."""
import sys, time
sys.path.insert(0, "tests"); import _util
import torch
dev = torch.device("cuda", 0)
feats = torch.randn(40, 1598, 80).pin_memory()
print("pinned:", feats.is_pinned(), "MB:", feats.numel() * 4 / 1e6)
def copy_ms(stream=None, n=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if stream is None:
            d = feats.to(dev, non_blocking=True)
        else:
            with torch.cuda.stream(stream):
                d = feats.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n
print("H2D same stream, idle GPU: %.3f ms" % copy_ms())
side = torch.cuda.Stream(dev)
print("H2D side stream, idle GPU: %.3f ms" % copy_ms(side))
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
def busy(n=40):
    for _ in range(n):
        torch.mm(a, a)
busy(5); torch.cuda.synchronize()
t0 = time.perf_counter(); busy(); torch.cuda.synchronize(); tb = (time.perf_counter() - t0) * 1e3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
with torch.cuda.stream(side):
    e0.record(side); d = feats.to(dev, non_blocking=True); e1.record(side)
busy(); torch.cuda.synchronize()
print("busy loop %.1f ms; H2D on the side stream next to it: %.3f ms" % (tb, e0.elapsed_time(e1)))
