#!/bin/bash
# ncu --set full capture of the batched-permanent walk: [KREGEX=laplace_walk] tools/ncu_batch.sh N B TAG
N=${1:-20}; B=${2:-2000}; TAG=${3:-batch}
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-perm_hyper} -c 1 -f -o gpurun_out/${TAG}_n${N} \
    python tools/batch_probe.py $N:$B > gpurun_out/${TAG}_n${N}.log 2>&1
ncu -i gpurun_out/${TAG}_n${N}.ncu-rep --page raw --csv > gpurun_out/${TAG}_n${N}_raw.csv 2>/dev/null
