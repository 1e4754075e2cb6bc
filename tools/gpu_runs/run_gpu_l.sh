set -x
timeout 900 python -m pytest tests/test_bev_pool_gpu.py -m gpu -x -q -k "cameras or lift" > gpurun_out/l_pytest.log 2>&1; tail -6 gpurun_out/l_pytest.log
for L in 3 5 7; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_final_l$L python tools/conv_v6_one.py $L > gpurun_out/ncu_final_l$L.log 2>&1; tail -1 gpurun_out/ncu_final_l$L.log
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/frame_once.py 3 > gpurun_out/frame_once.log 2>&1; tail -1 gpurun_out/frame_once.log
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -c 300 gpurun_out/bench_l.json; tail -5 gpurun_out/bench_l.err
