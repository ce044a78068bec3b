"""Top source lines by warp-stall samples for one kernel of an ncu report (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = {}; fpath = None; hdr = None; seen_fn = set(); first_fn = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; hdr = None
        if fpath in seen_fn: break          # the file list repeats for every captured instance: keep the first
        seen_fn.add(fpath); continue
    if r[0] == "Function Name":
        if first_fn is None: first_fn = r[1]
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    try: n = int(r[si] or 0); ni = int(r[ii] or 0)
    except ValueError: continue
    if not r[0]: continue        # SASS rows have no line number; keep the per-source-line rows only
    k = (fpath, r[0]); e = agg.setdefault(k, [0, 0, r[1].strip()[:110]]); e[0] += n; e[1] += ni
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print(f"{kern}: {tot} samples, {toti} warp-instructions over all captured instances")
for (f, ln), (n, ni, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/max(tot,1):5.1f}% smp {100*ni/max(toti,1):5.1f}% inst  {f}:{ln:>4}  {src}")
