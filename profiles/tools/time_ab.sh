# usage: VARIANTS="prev b200" bash profiles/tools/time_ab.sh  -- times phase-A-only (SimpleDocking3d) and the full C4 workload
for lib in ${VARIANTS:-b200}; do
  echo "== $lib"
  DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so python profiles/tools/time_scenarios.py 2>&1 | grep -E "SimpleDocking3d .*warp_rays|ObstaclesDocking3d .*warp_rays.*spheres"
done
