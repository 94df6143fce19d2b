#!/bin/bash
# Compile-time evidence for profiles/: ptxas -v (registers, spills) of every kernel
# family and a SASS excerpt of the headline instantiation.  No GPU needed.
#   tools/compile_evidence.sh r02
set -e
TAG=${1:-r02}
cd "$(dirname "$0")/.."
CS=piquasso_b200/csrc
OUT=profiles/${TAG}_ptxas_v.txt
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v"
: > $OUT
summarise() {  # stdin: nvcc -Xptxas -v output
    grep -E "Compiling entry function|registers|spill" | paste - - - | \
    sed -E "s/ptxas info    : Compiling entry function '([^']+)' for 'sm_100a'/\1/; s/ptxas info    : //g" | \
    while read -r line; do
        name=$(echo "$line" | awk '{print $1}' | c++filt 2>/dev/null || true)
        rest=$(echo "$line" | cut -d' ' -f2-)
        echo "$name | $rest"
    done
}
echo "# nvcc $FLAGS  (nvcc $(nvcc --version | grep release | sed 's/.*release //'))" >> $OUT
for part in "2 31 40"; do
    set -- $part
    echo "## pqperm_kernels_binary.cu part $1 (NC $2..$3)" >> $OUT
    nvcc $FLAGS -DPQ_BIN_PART=$1 -DPQ_BIN_LO=$2 -DPQ_BIN_HI=$3 -c $CS/pqperm_kernels_binary.cu -o /tmp/ev_bin$1.o 2>&1 | summarise >> $OUT
done
echo "## pqperm_kernels_laplace.cu (unit columns, leave-one-out sums: the sampler's kernels)" >> $OUT
nvcc $FLAGS -DPQ_LAP_UNIT=1 -DPQ_LAP_MODE=0 -c $CS/pqperm_kernels_laplace.cu -o /tmp/ev_lap.o 2>&1 | summarise >> $OUT
for part in 0 1 2; do
    echo "## pqperm_kernels_laplace.cu (unit columns, full product only: batched permanents, part $part)" >> $OUT
    nvcc $FLAGS -DPQ_LAP_UNIT=1 -DPQ_LAP_MODE=2 -DPQ_LAP_PART=$part -c $CS/pqperm_kernels_laplace.cu -o /tmp/ev_lapp.o 2>&1 | summarise >> $OUT
done
echo "## pqperm_kernels_permhyper.cu (batched permanents, hypercube flavour)" >> $OUT
nvcc $FLAGS -c $CS/pqperm_kernels_permhyper.cu -o /tmp/ev_hyp.o 2>&1 | summarise >> $OUT
echo "## pqperm_kernels_generic.cu" >> $OUT
nvcc $FLAGS -c $CS/pqperm_kernels_generic.cu -o /tmp/ev_gen.o 2>&1 | summarise >> $OUT
echo "## pqperm_arbiter.cu" >> $OUT
nvcc $FLAGS -fmad=false -c $CS/pqperm_arbiter.cu -o /tmp/ev_arb.o 2>&1 | summarise >> $OUT
# SASS of the headline kernel
SYM=$(cuobjdump -elf /tmp/ev_bin2.o 2>/dev/null | grep -o "_ZN6pqperm19perm_walk_binary_pmILi40ELi3ELi64EE[A-Za-z0-9_]*" | head -1)
cuobjdump -sass -fun "$SYM" /tmp/ev_bin2.o > /tmp/ev_n40.sass
S=profiles/${TAG}_sass_perm_walk_binary_pm_40_3_64.txt
{
  echo "# cuobjdump -sass of perm_walk_binary_pm<40,3,64> (sm_100a), $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/ev_n40.sass) instructions"
  echo "# opcode histogram:"
  grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9] )?[A-Z0-9_.]+" /tmp/ev_n40.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -24
  echo "# loops (tools/sass_loops.py):"
  python tools/sass_loops.py /tmp/ev_n40.sass
  echo "# first 160 instructions of the block loop (operands from the constant bank: LDCU / c[0x0][..]):"
  grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/ev_n40.sass | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | awk 'NR>=120 && NR<280'
} > $S
echo wrote $OUT $S
