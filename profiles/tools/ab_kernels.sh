# usage: VARIANTS="b200 a3" bash profiles/tools/ab_kernels.sh -- un-profiled step times, then per-kernel ncu durations (C4, split)
mkdir -p gpurun_out
for lib in ${VARIANTS:-b200}; do
  export DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so
  python profiles/tools/time_split.py 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_warp --launch-skip 200 -c 40 --csv --log-file gpurun_out/ab_$lib.csv python profiles/tools/run_scenario.py ObstaclesDocking3d split 3 125 > /dev/null 2>&1
  python profiles/tools/kernel_times.py gpurun_out/ab_$lib.csv
done
