timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitize_final.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/sanitize_final.log
