set -x
python tools/conv_v6_bench.py > gpurun_out/v6d.txt 2>&1; tail -11 gpurun_out/v6d.txt
BEVB200_V6_R=1 python tools/conv_v6_bench.py > gpurun_out/v6d_r1.txt 2>&1; tail -1 gpurun_out/v6d_r1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_v6d_s2 python tools/conv_v6_one.py 3 > gpurun_out/ncu_v6d_s2.log 2>&1; tail -2 gpurun_out/ncu_v6d_s2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_v6d_s4 python tools/conv_v6_one.py 7 > gpurun_out/ncu_v6d_s4.log 2>&1; tail -2 gpurun_out/ncu_v6d_s4.log
timeout 1200 python -m pytest tests/test_spconv_gpu.py tests/test_shims_gpu.py -m gpu -x -q > gpurun_out/v6d_pytest.log 2>&1; tail -8 gpurun_out/v6d_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 3000 gpurun_out/bench_d.json; tail -15 gpurun_out/bench_d.err
