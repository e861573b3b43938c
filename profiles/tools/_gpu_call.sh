mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r2_tests2.log 2>&1
tail -45 gpurun_out/r2_tests2.log
VARIANTS="b200 fold r112 c3 c5" bash profiles/tools/ab.sh
TAG=v2a bash profiles/tools/prof_pipe.sh
ls -la gpurun_out/*.ncu-rep
