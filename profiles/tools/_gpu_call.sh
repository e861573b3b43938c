mkdir -p gpurun_out
VARIANTS="b200 pdl2 npdl" bash profiles/tools/ab.sh
