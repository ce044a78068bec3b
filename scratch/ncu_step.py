"""Two full headline encoder steps (Type A Small N=2, 40 x 16 s, bf16 autocast) inside a cudaProfiler range: the ncu
launch-list target (ncu --profile-from-start off)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "h-net-mamba-asr_b200"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch
import bench, dcasr_b200 as dd
wl = sys.argv[1] if len(sys.argv) > 1 else "A_small_N2"
W = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
torch.manual_seed(1)
torch.backends.cudnn.benchmark = True
enc = dd.DCASREncoder(**W["kw"]).to(dev)
kw = W["kw"]
bench.set_routers(enc, kw["N"] if kw["arch_type"] == "A" else kw["N"] ** 0.5)
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
    return loss
for _ in range(4):
    step()
torch.cuda.synchronize(); torch.cuda.profiler.start()
for _ in range(int(os.environ.get("NCU_STEPS", "2"))):
    step()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok", float(step()))
