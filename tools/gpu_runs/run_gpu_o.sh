set -x
timeout 600 python -m pytest tests/test_spconv_gpu.py -m gpu -x -q -k "backward or gradient or autograd" > gpurun_out/o_pytest.log 2>&1; tail -15 gpurun_out/o_pytest.log
timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_tc.txt 2>&1; tail -12 gpurun_out/wgrad_tc.txt
BEVB200_WGRAD_TC=0 timeout 600 python tools/wgrad_bench.py > gpurun_out/wgrad_simt.txt 2>&1; tail -12 gpurun_out/wgrad_simt.txt
