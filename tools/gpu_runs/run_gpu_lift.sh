set -x
python tools/lift_bench.py C2 > gpurun_out/lift_c2.txt 2>&1; cat gpurun_out/lift_c2.txt
python tools/lift_bench.py C5 > gpurun_out/lift_c5.txt 2>&1; cat gpurun_out/lift_c5.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/lift_launches.csv python tools/lift_bench.py C2 > /dev/null 2>&1
grep -E "lift_|bevpool_fwd" gpurun_out/lift_launches.csv | awk -F'","' '{print $5, $13, $14, $15}' | tail -40
