"""Dev tool: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table of
one frame (the launches between two consecutive bevpool forward launches)."""
import csv, sys, collections, re

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv = H.index("Kernel Name"), H.index("Metric Value")
launches = [(r[kn], float(r[mv].replace(",", "")) / 1000.0) for r in rows[hdr + 1:] if r[mv].replace(",", "").replace(".", "").isdigit()]
marks = [i for i, (k, _) in enumerate(launches) if "bevpool_fwd" in k and "fixup" not in k]
a, b = marks[-2], marks[-1]
frame = launches[a:b]
agg = collections.OrderedDict()
for k, us in frame:
    k = re.sub(r"^void ", "", k); k = re.sub(r"bevb200::", "", k); k = re.sub(r"\(.*$", "", k)[:70]
    t, n = agg.get(k, (0.0, 0))
    agg[k] = (t + us, n + 1)
total = sum(t for t, _ in agg.values())
print("%d launches, %.1f us summed" % (len(frame), total))
print("| share | total us | launches | us/launch | kernel |\n|---|---|---|---|---|")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| %.1f%% | %.1f | %d | %.2f | `%s` |" % (100 * t / total, t, n, t / n, k))
