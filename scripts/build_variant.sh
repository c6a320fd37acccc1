#!/bin/bash
# Builds an alternative liba3vlm_b200 with extra -D flags for A/B tests: scripts/build_variant.sh NAME "-DX=1 ..."
# -> a3vlm_b200/lib/variants/NAME.so (select with A3_LIB_PATH)
set -e
cd "$(dirname "$0")/../a3vlm_b200/csrc"
NAME=$1; shift; FLAGS="$*"
mkdir -p ../../build/var_$NAME ../lib/variants
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -I. --expt-relaxed-constexpr $FLAGS -c $f -o ../../build/var_$NAME/${f%.cu}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/$NAME.so ../../build/var_$NAME/*.o -lcudart
echo built ../lib/variants/$NAME.so
