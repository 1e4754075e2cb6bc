timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference > gpurun_out/bench_c4g.json 2> gpurun_out/bench_c4g.err; tail -3 gpurun_out/bench_c4g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c4g.json').read().strip().splitlines()[-1])
print('value', d['value'], 'c4', {k:v for k,v in d['c4'].items() if k in ('frames_per_s','ms_per_frame','graph','stages_ms')})
PY
