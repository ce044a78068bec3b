"""fp32 (decode) ConvSubsampling4 forward at the headline shape: per-kernel device time (torch.profiler)."""
import sys
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
from torch.profiler import profile, ProfilerActivity
dev = "cuda"
torch.manual_seed(0)
sub = dd.ConvSubsampling4(80, 384).to(dev)
feats = torch.randn(40, 1598, 80, device=dev); lens = torch.full((40,), 1598, device=dev)
def step():
    with torch.no_grad():
        return sub(feats, lens)[0]
for _ in range(3): y = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y = step()
e1.record(); torch.cuda.synchronize()
print(f"fp32 front end forward: {e0.elapsed_time(e1) / 5:.2f} ms  checksum {float(y.double().abs().mean()):.8f}")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
    print(f"{e.device_time_total:9.0f} us  x{e.count:<3d} {e.key[:100]}")
