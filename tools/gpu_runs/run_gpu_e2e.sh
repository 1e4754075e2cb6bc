timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference --no-c4 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -3 gpurun_out/bench_e2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'mat', d['e2e_materialised']['value'], 'eager', d['eager'])
PY
