for L in 5 3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/prof_wg_l$L python tools/wgrad_one.py $L > gpurun_out/ncu_wg_l$L.log 2>&1; tail -1 gpurun_out/ncu_wg_l$L.log
done
