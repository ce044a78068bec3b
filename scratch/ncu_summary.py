"""Markdown summary of an `ncu --set full` report: one row per captured kernel launch with the metrics the roofline
discussion in DESIGN.md uses."""
import csv, subprocess, sys, re
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k, f=1.0):
    try: return float(r[idx[k]]) * f
    except Exception: return float("nan")
cols = [("dur us", "gpu__time_duration.sum", 1), ("DRAM rd MB", "dram__bytes_read.sum", 1), ("DRAM wr MB", "dram__bytes_write.sum", 1),
        ("DRAM %pk", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
        ("tensor %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("regs", "launch__registers_per_thread", 1),
        ("inst M", "smsp__inst_executed.sum", 1e-6),
        ("st barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1),
        ("st long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
        ("st short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1)]
units = rows[1]
lines = [f"# ncu --set full summary: `{rep.split('/')[-1]}`", "",
         "Command: `ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:\"ssd_|conv_|gemm_bf16|norm_|pack_mixer\" python scratch/ncu_target.py`",
         "(one bidirectional MambaBlock forward+backward at the headline shape B = 40 x 398 frames, d = 384, bf16, plus one",
         "chunk/dechunk round trip).  Per-launch values; durations are cold-cache and serialised (ncu replays each kernel).",
         "`st *` = average warps stalled for that reason per issue-active cycle.", "",
         "| kernel | " + " | ".join(c[0] for c in cols) + " |", "|---|" + "---|" * len(cols)]
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "").replace("hnb::", "")
    name = re.sub(r"<unnamed>::|<|>", " ", name).strip()[:46]
    vals = []
    for label, key, f in cols:
        v = g(r, key, f)
        if label.endswith("MB") and key in idx and units[idx[key]] == "Kbyte": v /= 1e3
        if label.endswith("MB") and key in idx and units[idx[key]] == "byte": v /= 1e6
        if label == "dur us" and key in idx and units[idx[key]] in ("ns", "nsecond"): v /= 1e3
        if label == "dur us" and key in idx and units[idx[key]] in ("ms", "msecond"): v *= 1e3
        vals.append("n/a" if v != v else (f"{v:.0f}" if abs(v) >= 100 else f"{v:.1f}"))
    lines.append(f"| {name} | " + " | ".join(vals) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-40:]))
