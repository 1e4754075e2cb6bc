set -x
python tools/conv_v6_bench.py > gpurun_out/v6h.txt 2>&1; tail -11 gpurun_out/v6h.txt
BEVB200_V6_NSB=2 python tools/conv_v6_bench.py > gpurun_out/v6h_nsb2.txt 2>&1; tail -11 gpurun_out/v6h_nsb2.txt
timeout 1200 python -m pytest tests/test_bev_pool_gpu.py -m gpu -x -q > gpurun_out/v6h_pytest.log 2>&1; tail -12 gpurun_out/v6h_pytest.log
timeout 1500 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -c 600 gpurun_out/bench_h.json; tail -15 gpurun_out/bench_h.err
