#!/bin/bash
# Diagnostic build of the library with clock64 stamps in the attention kernel (-DBLADE_TRACE); use it with
#   BLADE_ASA_LIB=video_blade_b200/lib/libblade_asa_trace.so python tools/trace_attn.py [wan|cog]
set -e
cd "$(dirname "$0")/.."
C=video_blade_b200/csrc; L=video_blade_b200/lib; mkdir -p $L/trace
for s in capi mask_kernels attn_kernel estimator_kernel; do
  nvcc -c $C/$s.cu -o $L/trace/$s.o -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
       --expt-relaxed-constexpr -I include -DBLADE_TRACE &
done
wait
nvcc -shared -o $L/libblade_asa_trace.so $L/trace/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo $L/libblade_asa_trace.so
