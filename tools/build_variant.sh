#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>
# -> shinestacker_b200/variants/lib_<name>.so (select it with SHINESTACK_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../shinestacker_b200/csrc"
name=$1; shift
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
     -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -shared "$@" -o ../variants/lib_$name.so ss_lib.cu
echo built ../variants/lib_$name.so
