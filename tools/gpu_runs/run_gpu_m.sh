set -x
timeout 300 ./tools/bin/gather4_probe > gpurun_out/gather4_probe.txt 2>&1; tail -80 gpurun_out/gather4_probe.txt
timeout 300 python tools/conv_v6_bench.py small > gpurun_out/v6tma_small.txt 2>&1; tail -12 gpurun_out/v6tma_small.txt
timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6tma_bench.txt 2>&1; tail -12 gpurun_out/v6tma_bench.txt
BEVB200_V6_TMA=0 timeout 600 python tools/conv_v6_bench.py > gpurun_out/v6ldgsts_bench.txt 2>&1; tail -3 gpurun_out/v6ldgsts_bench.txt
