set -x
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 600 gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
