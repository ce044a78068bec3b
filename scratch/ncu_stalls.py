"""Summarise warp-stall samples of one kernel from an ncu report: top SASS instructions and per-reason totals."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr): break          # next kernel instance
    body.append(r)
si = hdr.index("# Samples")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print(f"{kern}: {len(body)} SASS instructions, {tot} samples")
agg = {h: sum(int(r[hdr.index(h)] or 0) for r in body) for h in reasons}
print("by reason:", {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]})
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top]:
    rs = sorted(((int(r[hdr.index(h)] or 0), h) for h in reasons), reverse=True)[:2]
    print(f"  {100*int(r[si])/tot:5.1f}%  {r[1][:70]:70s} {rs}")
