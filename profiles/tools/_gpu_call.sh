mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r2_tests11.log 2>&1
tail -8 gpurun_out/r2_tests11.log
VARIANTS="b200 w2" bash profiles/tools/ab.sh
TAG=v2i bash profiles/tools/prof_pipe.sh
