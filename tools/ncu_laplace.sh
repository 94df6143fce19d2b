#!/bin/bash
# ncu --set full capture of the Laplace walk on a k-column batch: tools/ncu_laplace.sh K BATCH TAG
K=${1:-24}; B=${2:-128}; TAG=${3:-lap}
ncu --set full --clock-control none --import-source on -k regex:laplace_walk -c 1 -f -o gpurun_out/${TAG}_k${K} \
    python tools/run_laplace.py $K $B > gpurun_out/${TAG}_k${K}.log 2>&1
ncu -i gpurun_out/${TAG}_k${K}.ncu-rep --page raw --csv > gpurun_out/${TAG}_k${K}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k${K}.ncu-rep --page source --csv > gpurun_out/${TAG}_k${K}_src.csv 2>/dev/null
