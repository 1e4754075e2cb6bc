timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/tests_final3.log 2>&1; tail -4 gpurun_out/tests_final3.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/frame_once.py 3 > gpurun_out/frame_once.log 2>&1; tail -1 gpurun_out/frame_once.log
timeout 1500 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; tail -3 gpurun_out/bench_final3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final3.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'stages', d['stages_ms'], 'roofline', d['roofline']['frac'], d['roofline']['ms'], 'training', d['training']['spconv_bwd_ms_21_convs'], 'c4', d['c4']['frames_per_s'], d['c4'].get('graph'))
PY
