set -x
python tools/conv_v6_bench.py > gpurun_out/v6g.txt 2>&1; tail -11 gpurun_out/v6g.txt
BEVB200_V6_CTAS=1 python tools/conv_v6_bench.py > gpurun_out/v6g_c1.txt 2>&1; tail -11 gpurun_out/v6g_c1.txt
BEVB200_V6_LAG=0 python tools/conv_v6_bench.py > gpurun_out/v6g_lag0.txt 2>&1; tail -11 gpurun_out/v6g_lag0.txt
BEVB200_V6_CTAS=1 BEVB200_V6_LAG=0 python tools/conv_v6_bench.py > gpurun_out/v6g_c1_lag0.txt 2>&1; tail -1 gpurun_out/v6g_c1_lag0.txt
BEVB200_V6_NSB=2 python tools/conv_v6_bench.py > gpurun_out/v6g_nsb2.txt 2>&1; tail -1 gpurun_out/v6g_nsb2.txt
