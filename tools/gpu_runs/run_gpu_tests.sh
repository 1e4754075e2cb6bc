timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/tests_final.log 2>&1; tail -6 gpurun_out/tests_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
