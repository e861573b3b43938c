# the evidence run of a round: plain bench (all configs), ncu launch list of the same command, full ncu capture of one step
TAG=${TAG:-v2}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
for c in C2 C3; do python bench.py --config $c --steps 400 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2>> gpurun_out/${TAG}_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 --burn-in 64 > gpurun_out/${TAG}_ncu_launches.log 2>&1
TAG=$TAG bash profiles/tools/prof_pipe.sh
for f in gpurun_out/${TAG}_bench_1gpu.json gpurun_out/${TAG}_bench_C2.json gpurun_out/${TAG}_bench_C3.json gpurun_out/${TAG}_bench_reference.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d.get('roofline',{}).get('step',{}).get('launches_ms'))
"; done
