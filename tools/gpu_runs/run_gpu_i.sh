set -x
timeout 900 python -m pytest tests/test_bev_pool_gpu.py -m gpu -x -q -k "lift" > gpurun_out/i_pytest.log 2>&1; tail -5 gpurun_out/i_pytest.log
python tools/lift_bench.py C2 > gpurun_out/lift_c2.txt 2>&1; cat gpurun_out/lift_c2.txt
python tools/lift_bench.py C5 > gpurun_out/lift_c5.txt 2>&1; cat gpurun_out/lift_c5.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/frame_once.py 3 > gpurun_out/frame_once.log 2>&1; tail -2 gpurun_out/frame_once.log
wc -l gpurun_out/launches_r2.csv
