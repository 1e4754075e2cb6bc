timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6l_skip.txt 2>&1; tail -11 gpurun_out/v6l_skip.txt
BEVB200_V6_SKIPZERO=0 timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6l_noskip.txt 2>&1; tail -3 gpurun_out/v6l_noskip.txt
