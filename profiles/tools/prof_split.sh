# full ncu capture of the two launches of the split layout in steady state (C4): gpurun_out/prof_split_$TAG.ncu-rep
set -x
TAG=${TAG:-cur}
python profiles/tools/run_scenario.py ObstaclesDocking3d split 3 135 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel --launch-skip 264 --launch-count 2 -f -o gpurun_out/prof_split_$TAG python profiles/tools/run_scenario.py ObstaclesDocking3d split 3 135 > gpurun_out/prof_split_$TAG.log 2>&1
