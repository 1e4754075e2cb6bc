set -x
python tools/conv_v6_bench.py > gpurun_out/v6j.txt 2>&1; tail -11 gpurun_out/v6j.txt | cut -c1-190
BEVB200_V6_LAG=2 python tools/conv_v6_bench.py > gpurun_out/v6j_lag2.txt 2>&1; tail -11 gpurun_out/v6j_lag2.txt | cut -c1-190
BEVB200_V6_GFENCE=0 python tools/conv_v6_bench.py > gpurun_out/v6j_nofence.txt 2>&1; tail -11 gpurun_out/v6j_nofence.txt | cut -c1-190
BEVB200_V6_GFENCE=0 BEVB200_V6_LAG=2 BEVB200_V6_NSB=2 python tools/conv_v6_bench.py > gpurun_out/v6j_all.txt 2>&1; tail -11 gpurun_out/v6j_all.txt | cut -c1-190
timeout 900 python -m pytest tests/test_spconv_gpu.py -m gpu -x -q -k "native_plan or weight_gradient" > gpurun_out/j_pytest.log 2>&1; tail -5 gpurun_out/j_pytest.log
