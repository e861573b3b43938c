# usage: VARIANTS="b200 a3 a5" bash profiles/tools/ab.sh [--quick]   (variants = gym_dockauv_b200/_lib/libdockauv_<tag>.so)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in ${VARIANTS:-b200}; do
  DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so python profiles/tools/ab_step.py $lib $1 2>&1 | tail -1 | tee -a gpurun_out/ab_step.log
done
done
