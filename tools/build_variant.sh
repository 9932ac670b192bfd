#!/bin/bash
# A/B build of the library: recompiles attn_kernel.cu with extra -D flags, reuses the other objects of the regular build.
#   tools/build_variant.sh poly1 -DBLADE_POLY_PAIRS=1   ->  video_blade_b200/lib/variants/libblade_asa_poly1.so
# use with BLADE_ASA_LIB=<that .so> (tools/ab_lib.sh)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=video_blade_b200/csrc; L=video_blade_b200/lib; mkdir -p $L/variants
nvcc -c $C/attn_kernel.cu -o $L/variants/attn_kernel_$name.o -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
     -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v -I include "$@" 2> $L/variants/ptxas_$name.log
objs=$(ls $L/*.o | grep -v attn_kernel.o)
nvcc -shared -o $L/variants/libblade_asa_$name.so $objs $L/variants/attn_kernel_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static
grep -A2 "asa_attn_kernelILi128ELb1ELb1\|asa_attn_kernelILi64ELb1ELb1" $L/variants/ptxas_$name.log | grep spill
echo $L/variants/libblade_asa_$name.so
