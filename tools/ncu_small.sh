#!/bin/bash
# ncu --set full captures of the generic walk on the small BASELINE shapes:
#   tools/ncu_small.sh  ->  gpurun_out/small_{cfg3,n20}_raw.csv
# (-s skips the warm-up launches so that a warm launch is captured)
ncu --set full --clock-control none --import-source on -k regex:perm_walk_generic -s 20 -c 1 -f -o gpurun_out/small_cfg3 \
    python tools/time_cfg3.py > gpurun_out/small_cfg3.log 2>&1
ncu -i gpurun_out/small_cfg3.ncu-rep --page raw --csv > gpurun_out/small_cfg3_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:perm_walk_generic -s 20 -c 1 -f -o gpurun_out/small_n20 \
    python tools/run_one.py 20 0 40 > gpurun_out/small_n20.log 2>&1
ncu -i gpurun_out/small_n20.ncu-rep --page raw --csv > gpurun_out/small_n20_raw.csv 2>/dev/null
