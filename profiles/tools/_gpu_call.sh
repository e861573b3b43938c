mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r2_tests23.log 2>&1
tail -12 gpurun_out/r2_tests23.log
VARIANTS="b200" bash profiles/tools/ab.sh
