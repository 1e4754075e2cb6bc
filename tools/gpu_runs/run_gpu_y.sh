timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6k_skip.txt 2>&1; tail -11 gpurun_out/v6k_skip.txt
BEVB200_V6_SKIPZERO=0 timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6k_noskip.txt 2>&1; tail -3 gpurun_out/v6k_noskip.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/y_pytest.log 2>&1; tail -8 gpurun_out/y_pytest.log
