"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum of one bench step) ->
profiles/<out>.json: DRAM traffic and duration per C-ABI entry point (sums over the kernels an entry point launches),
per step and per call.  bench.py reads it for the `traffic` field of its roofline object."""
import csv, json, re, sys, collections
src, out, steps = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
ENTRY = [("ssd_bwd", r"ssd_bwd_(dstate|dx|dbc|fused)"), ("ssd_fwd", r"ssd_fwd2?_tc|ssd_tables"), ("gemm_bf16", r"gemm_bf16_(pair_)?kernel"),
         ("conv_bwd", r"conv_bwd_kernel"), ("conv_fwd", r"conv_fwd_kernel"), ("gated_norm_bwd", r"gated_norm_bwd"),
         ("gated_norm_fwd", r"gated_norm_fwd"), ("layernorm_bwd", r"layernorm_bwd"), ("layernorm_fwd", r"layernorm_fwd"),
         ("pack_mixer_params", r"pack_mixer"), ("subsample_conv1_fwd", r"sub_conv1_fwd"), ("subsample_conv1_bwd", r"sub_conv1_bwd")]
CALLS = {"ssd_bwd": 2, "ssd_fwd": 2}                       # kernels per C call (state-gradient pass + fused kernel; tables + forward)
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r); i0 = rows.index(hdr)
kn, mn, mv, mu, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.defaultdict(lambda: {"kernels": 0, "dram_bytes": 0.0, "us": 0.0})
seen = set()
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
for r in rows[i0 + 1:]:
    name = r[kn]
    ent = next((e for e, pat in ENTRY if re.search(pat, name)), None)
    if ent is None: continue
    v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1)
    if r[mn].startswith("dram__bytes"): per[ent]["dram_bytes"] += v
    elif r[mn].startswith("gpu__time"): per[ent]["us"] += v
    if (r[idc], ent) not in seen: seen.add((r[idc], ent)); per[ent]["kernels"] += 1
res = {}
for e, d in per.items():
    calls = d["kernels"] / CALLS.get(e, 1)
    res[e] = {"calls_captured": calls, "dram_bytes_per_call": d["dram_bytes"] / max(calls, 1), "us_per_call_under_ncu": d["us"] / max(calls, 1),
              "calls_per_step": calls / steps}
json.dump({"source": src, "steps_captured": steps, "workload": "A_small_N2", "entries": res}, open(out, "w"), indent=1)
for e, d in sorted(res.items(), key=lambda kv: -kv[1]["us_per_call_under_ncu"] * kv[1]["calls_captured"]):
    print(f"{e:22s} calls {d['calls_captured']:6.1f}  {d['dram_bytes_per_call']/1e6:9.1f} MB/call  {d['us_per_call_under_ncu']:8.1f} us/call")
