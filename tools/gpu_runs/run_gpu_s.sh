set -x
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -x -q -k "backward or gradient or autograd" > gpurun_out/s_pytest.log 2>&1; tail -8 gpurun_out/s_pytest.log
timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_tc3.txt 2>&1; tail -12 gpurun_out/wgrad_tc3.txt
set +x
for A in 0 1 2 3; do
echo "== ablate $A"
BEVB200_WGRAD_ABLATE=$A timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad_tc_kernel" --csv --log-file gpurun_out/wg3_abl$A.csv python tools/wgrad_bench.py > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/wg3_abl$A.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki][:44], float(r[vi].replace(',', ''))/1e3) for r in rows[hdr + 1:] if len(r) > vi]
print(" ".join("%.0f" % v for n, v in seq[7::8]))
PY
done
