"""Dev tool: summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
log into a markdown table of one frame (the launches between two consecutive bevpool forward launches)."""
import csv, sys, collections, re

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kid, kn, mn, mv = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try:
        val = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[H.index("Metric Unit")] if "Metric Unit" in H else ""
    d = launch.setdefault(r[kid], {"name": r[kn]})
    if r[mn] == "gpu__time_duration.sum":
        d["us"] = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit.startswith("us") else val / 1000.0)
    elif r[mn].startswith("dram__bytes"):
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
        d["dram"] = d.get("dram", 0.0) + val * scale
launches = list(launch.values())
marks = [i for i, d in enumerate(launches) if "bevpool_fwd" in d["name"] and "fixup" not in d["name"]]
frame = launches[marks[-2]:marks[-1]]
agg = collections.OrderedDict()
for d in frame:
    k = re.sub(r"^void ", "", d["name"]); k = re.sub(r"bevb200::", "", k); k = re.sub(r"\(.*$", "", k)[:70]
    t, n, b = agg.get(k, (0.0, 0, 0.0))
    agg[k] = (t + d.get("us", 0.0), n + 1, b + d.get("dram", 0.0))
total = sum(t for t, _, _ in agg.values())
print("%d launches, %.1f us summed" % (len(frame), total))
print("| share | total us | launches | us/launch | DRAM MB (rd+wr) | kernel |\n|---|---|---|---|---|---|")
for k, (t, n, b) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| %.1f%% | %.1f | %d | %.2f | %.1f | `%s` |" % (100 * t / total, t, n, t / n, b / 1e6, k))
