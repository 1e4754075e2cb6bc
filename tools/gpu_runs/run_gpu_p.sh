set -x
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -x -q -k "backward or gradient or autograd" > gpurun_out/p_pytest.log 2>&1; tail -15 gpurun_out/p_pytest.log
timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_tc2.txt 2>&1; tail -12 gpurun_out/wgrad_tc2.txt
BEVB200_WGRAD_ROWS=64 timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_tc2_r64.txt 2>&1; tail -12 gpurun_out/wgrad_tc2_r64.txt
BEVB200_WGRAD_ROWS=128 timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_tc2_r128.txt 2>&1; tail -12 gpurun_out/wgrad_tc2_r128.txt
