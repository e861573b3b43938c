# the evidence run of a round: plain bench, ncu launch list of the same command, full ncu capture of one step's launches
TAG=${TAG:-v9}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 --burn-in 64 > gpurun_out/${TAG}_ncu_launches.log 2>&1
TAG=$TAG bash profiles/tools/prof_pipe.sh
cat gpurun_out/${TAG}_bench_1gpu.json
