mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r2_tests13.log 2>&1
tail -4 gpurun_out/r2_tests13.log
VARIANTS="b200 cp0 dp0 tp0" bash profiles/tools/ab.sh
