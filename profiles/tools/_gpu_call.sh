mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_tests6.log 2>&1
tail -6 gpurun_out/r2_tests6.log
TAG=v2e bash profiles/tools/round_profile.sh
