timeout 1200 python -m pytest tests/test_spconv_gpu.py -m gpu -q -x > gpurun_out/pdl_pytest.log 2>&1; tail -3 gpurun_out/pdl_pytest.log
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference --no-c4 > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; tail -3 gpurun_out/bench_pdl.err
BEVB200_V6_PDL=0 timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference --no-c4 > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; tail -3 gpurun_out/bench_nopdl.err
python - <<'PY'
import json
for f in ('bench_pdl','bench_nopdl'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], d['ms_per_step'], 'stages', d['stages_ms'], 'e2e', d['e2e']['value'], 'eager', d['eager']['value'], 'conv', d['roofline']['ms'])
PY
