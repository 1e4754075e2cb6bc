set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -5 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
for L in 3 5 7; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_final_l$L python tools/conv_v6_one.py $L > gpurun_out/ncu_final_l$L.log 2>&1; tail -1 gpurun_out/ncu_final_l$L.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/prof_final_wg7 python tools/wgrad_one.py 7 > gpurun_out/ncu_final_wg7.log 2>&1; tail -1 gpurun_out/ncu_final_wg7.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/frame_once.py 3 > gpurun_out/frame_once.log 2>&1; tail -1 gpurun_out/frame_once.log
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json; tail -5 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 300 gpurun_out/bench_final_ref.json
