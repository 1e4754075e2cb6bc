timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; tail -3 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_final.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'c4', d['c4']['frames_per_s'], d['c4'].get('graph'))
PY
