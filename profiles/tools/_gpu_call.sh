for rep in 1 2; do for lib in b200 pm16 pm18 pm19; do DOCKAUV_LIB=$PWD/gym_dockauv_b200/_lib/libdockauv_$lib.so python profiles/tools/small_batches.py $lib 2>&1 | tail -1; done; done
