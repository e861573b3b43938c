mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --durations=3 ) > gpurun_out/r2_tests26.log 2>&1
tail -3 gpurun_out/r2_tests26.log
TAG=v2p bash profiles/tools/round_profile.sh
