set -x
timeout 180 python tools/conv_v6_bench.py small > gpurun_out/v6_small.txt 2>&1 || { echo SMALL_FAILED; tail -30 gpurun_out/v6_small.txt; exit 1; }
tail -12 gpurun_out/v6_small.txt
timeout 300 python tools/conv_v6_bench.py > gpurun_out/v6_bench.txt 2>&1; tail -12 gpurun_out/v6_bench.txt
BEVB200_V6_R=1 timeout 300 python tools/conv_v6_bench.py > gpurun_out/v6_bench_r1.txt 2>&1; tail -11 gpurun_out/v6_bench_r1.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/conv_v6_bench.py small > gpurun_out/v6_memcheck.log 2>&1; tail -5 gpurun_out/v6_memcheck.log
timeout 600 python tools/tc_check.py > gpurun_out/tc_check_v6.txt 2>&1; grep "prec 3" gpurun_out/tc_check_v6.txt
timeout 1200 python -m pytest tests/test_spconv_gpu.py -m gpu -x -q > gpurun_out/v6_pytest.log 2>&1; tail -8 gpurun_out/v6_pytest.log
