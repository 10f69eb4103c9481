#!/bin/bash
# Kernel-variant builds of the fused RK45 step kernel (development aid): oc_hjb.cu and oc_hjb_dist.cu are recompiled with the
# given -D flags and linked with the stock objects into optimal_crowds_b200/variants/liboc_<tag>.so.
# usage: scripts/build_variants.sh tag "-DOC_BX=192 -DOC_CTAS=2" [tag2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
C=optimal_crowds_b200/csrc
mkdir -p optimal_crowds_b200/variants
python -m optimal_crowds_b200.build > /dev/null
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fno-fast-math \
       $flags -Xptxas -v -c $C/oc_hjb.cu -o optimal_crowds_b200/variants/oc_hjb_$tag.o 2> optimal_crowds_b200/variants/ptxas_$tag.log
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fno-fast-math \
       $flags -c $C/oc_hjb_dist.cu -o optimal_crowds_b200/variants/oc_hjb_dist_$tag.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o optimal_crowds_b200/variants/liboc_$tag.so \
       $C/oc_api.o $C/oc_gcfm.o optimal_crowds_b200/variants/oc_hjb_$tag.o optimal_crowds_b200/variants/oc_hjb_dist_$tag.o -ldl
  echo "$tag: $(grep -A1 'hjb_fused_kernelILi3' optimal_crowds_b200/variants/ptxas_$tag.log | grep -o 'Used [0-9]* registers.*' | head -1) spills: $(grep -A1 'hjb_fused_kernelILi3' optimal_crowds_b200/variants/ptxas_$tag.log | grep -o '[0-9]* bytes spill stores' | head -1)"
done
