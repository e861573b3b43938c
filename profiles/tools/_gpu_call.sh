mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r2_tests27.log 2>&1
tail -4 gpurun_out/r2_tests27.log
VARIANTS="b200" bash profiles/tools/ab.sh --quick
