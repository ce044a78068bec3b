"""ncu launch-list CSV (gpu__time_duration.sum per launch) -> per-kernel totals and shares (markdown)."""
import csv, collections, re, sys
src = sys.argv[1]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r); i0 = rows.index(hdr)
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mn = hdr.index("Metric Name")
agg = collections.OrderedDict(); tot = 0.0; n = 0
for r in rows[i0 + 1:]:
    if not r[mn].startswith("gpu__time_duration"): continue     # (the CSV may carry DRAM byte counters too)
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] in ("ns", "nsecond") else (v * 1000 if r[mu] in ("ms", "msecond") else v)
    name = r[kn].replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"hnb::(<unnamed>::|\(anonymous namespace\)::)?", "hnb::", name)
    name = re.sub(r"<.*", "", name)[:64]
    e = agg.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += v; tot += v; n += 1
ours = sum(t for k, (c, t) in agg.items() if k.startswith("hnb::"))
print(f"captured launches: {n}; device time {tot/1000:.1f} ms; hand-written `hnb::` kernels {100*ours/tot:.1f} % of it\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 34]:
    print(f"| {k} | {c} | {t:.0f} | {100*t/tot:.1f}% |")
