mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --config C2 --steps 400 --no-cpu-baseline > gpurun_out/c2_try.json 2> gpurun_out/c2_try.err; python -c "
import json
d=json.loads(open('gpurun_out/c2_try.json').read().strip().splitlines()[-1]); print('C2', d['value'], d['ms_per_step'], d['roofline']['step']['launches_ms'])"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/c4_try.json 2> gpurun_out/c4_try.err; python -c "
import json
d=json.loads(open('gpurun_out/c4_try.json').read().strip().splitlines()[-1]); print('C4', d['value'], d['ms_per_step'], d['roofline']['traffic'], d['roofline']['fp64']['executed_pipe_frac'], d['roofline']['step']['traffic'])"
