#!/bin/bash
# Build an experiment variant of libpqperm.so from the working tree:
#   tools/build_variant.sh NAME "-DPQ_TRACE=1 ..."  ->  variants/libpqperm_NAME.so
# (load it with PQ_LIB_PATH=variants/libpqperm_NAME.so; the production library is untouched)
set -e
NAME=$1; FLAGS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=/tmp/pqvariant_$NAME
rm -rf $TMP; mkdir -p $TMP
cp -r $ROOT/piquasso_b200 $ROOT/include $TMP/
rm -rf $TMP/piquasso_b200/csrc/build $TMP/piquasso_b200/libpqperm.so
(cd $TMP && PQ_EXTRA_NVCC_FLAGS="$FLAGS" python -c "from piquasso_b200 import build; build.build(verbose=True)")
mkdir -p $ROOT/variants
cp $TMP/piquasso_b200/libpqperm.so $ROOT/variants/libpqperm_$NAME.so
echo built $ROOT/variants/libpqperm_$NAME.so
