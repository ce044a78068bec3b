"""Executed warp-instructions per SASS opcode for one kernel of an ncu report."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); si = hdr.index("Source")
cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(hdr): break
    toks = r[si].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    cnt[op.rstrip(";")] += int(r[ii] or 0)
tot = sum(cnt.values())
print(kern, "warp-instructions", tot)
for op, n in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30): print(f"  {100*n/tot:5.1f}%  {n:>10}  {op}")
