set -x
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/s3_plain.json 2>gpurun_out/s3_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/s3_ncu1.log 2>&1
python profiles/tools/run_scenario.py ObstaclesDocking3d split 3 135 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel --launch-skip 264 --launch-count 2 -f -o gpurun_out/prof_split_v7 python profiles/tools/run_scenario.py ObstaclesDocking3d split 3 135 > gpurun_out/s3_ncu2.log 2>&1
ls -la gpurun_out/
