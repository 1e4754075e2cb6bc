set -x
python tools/conv_v6_bench.py > gpurun_out/v6e.txt 2>&1; tail -11 gpurun_out/v6e.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spconv_v6_kernel -s 2 -c 1 -o gpurun_out/prof_v6e_s2 python tools/conv_v6_one.py 3 > gpurun_out/ncu_v6e_s2.log 2>&1; tail -1 gpurun_out/ncu_v6e_s2.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/v6e_pytest.log 2>&1; tail -8 gpurun_out/v6e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v6e_smoke.log 2>&1; tail -2 gpurun_out/v6e_smoke.log
timeout 1500 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -c 1500 gpurun_out/bench_e.json; tail -15 gpurun_out/bench_e.err
