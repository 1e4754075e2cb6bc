set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad|split_rows|spconv_v6_kernel|transpose|pack" --csv --log-file gpurun_out/wgrad_launches.csv python tools/wgrad_bench.py > gpurun_out/wgrad_ncu.txt 2>&1; tail -3 gpurun_out/wgrad_ncu.txt
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/wgrad_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki][:60], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
# print the last pass over each layer: the launches come in a fixed order per sparse_conv_backward call
for name, v in seq[-400:]:
    pass
agg = collections.OrderedDict()
for name, v in seq:
    agg.setdefault(name, []).append(v)
for name, vs in agg.items():
    print("%-62s n %4d  mean %9.1f us  min %9.1f  max %9.1f" % (name, len(vs), sum(vs) / len(vs) / 1e3, min(vs) / 1e3, max(vs) / 1e3))
# per-call sequence of the LAST 9 x (calls) -- show last 120 launches
print("---- last 150 launches")
for name, v in seq[-150:]:
    print("%-62s %9.1f us" % (name, v / 1e3))
PY
