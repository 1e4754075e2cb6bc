timeout 300 ./tools/bin/gather4_probe > gpurun_out/gather4_probe2.txt 2>&1; cat gpurun_out/gather4_probe2.txt
