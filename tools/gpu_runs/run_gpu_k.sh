set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; tail -6 gpurun_out/k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; tail -2 gpurun_out/k_smoke.log
timeout 1500 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -c 400 gpurun_out/bench_k.json; tail -15 gpurun_out/bench_k.err
