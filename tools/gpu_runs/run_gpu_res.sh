timeout 1200 python -m pytest tests/test_spconv_gpu.py -m gpu -q -k "native or encoder or plan or parity or full" > gpurun_out/res_pytest.log 2>&1; tail -5 gpurun_out/res_pytest.log
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference --no-c4 > gpurun_out/bench_res.json 2> gpurun_out/bench_res.err; tail -3 gpurun_out/bench_res.err
BEVB200_ENCODER_F32_RESIDUAL=1 timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 --no-gpu-reference --no-c4 > gpurun_out/bench_res_f32.json 2> gpurun_out/bench_res_f32.err; tail -3 gpurun_out/bench_res_f32.err
python - <<'PY'
import json
for f in ('bench_res','bench_res_f32'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], d['ms_per_step'], 'stages', d['stages_ms'], 'e2e', d['e2e']['value'])
PY
