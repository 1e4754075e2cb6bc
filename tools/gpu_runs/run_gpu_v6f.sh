set -x
python tools/conv_v6_bench.py > gpurun_out/v6f.txt 2>&1; tail -11 gpurun_out/v6f.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_v6f_s2 python tools/conv_v6_one.py 3 > gpurun_out/ncu_v6f_s2.log 2>&1; tail -1 gpurun_out/ncu_v6f_s2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_v6f_s4 python tools/conv_v6_one.py 7 > gpurun_out/ncu_v6f_s4.log 2>&1; tail -1 gpurun_out/ncu_v6f_s4.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/v6f_pytest.log 2>&1; tail -8 gpurun_out/v6f_pytest.log
