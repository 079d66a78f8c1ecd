#!/bin/bash
# profiles/build_variant.sh NAME [extra nvcc flags...]: libzfista_b200 with the batched-kernel
# translation units (zf_batched*.cu) compiled with the extra flags (the other objects come from
# the regular in-tree build) -> profiles/variants/libzf_NAME.so; select it with ZFISTA_B200_LIB=...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
python -m zfista_b200.build > /dev/null
mkdir -p profiles/variants/$name
for f in zfista_b200/csrc/zf_batched*.cu; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    -Iinclude -Izfista_b200/csrc "$@" -c $f -o profiles/variants/$name/$(basename $f .cu).o &
done
wait
objs=$(ls zfista_b200/build/*.o | grep -v zf_batched)
nvcc -shared -o profiles/variants/libzf_$name.so profiles/variants/$name/*.o $objs -lcudart 2>/dev/null
echo profiles/variants/libzf_$name.so
