"""Which ATen kernels still launch inside one headline step (torch.profiler, CUDA activity), grouped by op."""
import os, sys, collections
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "h-net-mamba-asr_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch
import bench, dcasr_b200 as dd
from torch.profiler import profile, ProfilerActivity
W = bench.WORKLOADS["A_small_N2"]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
torch.backends.cudnn.benchmark = True
enc = dd.DCASREncoder(**W["kw"]).to(dev)
bench.set_routers(enc, 2)
f, l = bench.synth_batch(W["batch"], W["seconds"], 1)
f, l = f.to(dev), l.to(dev)
params = list(enc.parameters())
def step():
    for p in params:
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc(f, l)
    loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
    loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_time_total > 0 and e.key.startswith("aten::"):
        rows.append((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"aten ops with device time: {sum(r[1] for r in rows)} calls, {tot:.0f} us self device time")
for t, c, k, sh in rows[:45]:
    if t > 0:
        print(f"{t:8.0f} us  x{c:<3d} {k:34s} {sh}")
