for lib in ${VARIANTS:-b200}; do
  DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '%.4g'%d['value'], '%.3f ms'%d['ms_per_step'], 'pipe %.3f'%d['pipe']['frac'])"
done
