mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:rays_thread" --launch-skip 264 --launch-count 1 -f -o gpurun_out/prof_rays_sp python profiles/tools/run_scenario.py ObstaclesDocking3d pipeline 3 135 > gpurun_out/prof_rays_sp.log 2>&1
tail -2 gpurun_out/prof_rays_sp.log
