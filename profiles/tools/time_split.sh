for lib in ${VARIANTS:-b200}; do DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so python profiles/tools/time_split.py 2>&1 | tail -1; done
