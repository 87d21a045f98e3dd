#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...] -- a second build of libimd_b200.so whose force kernels (forces.cu, all four
# instances) are compiled with extra flags, for kernel experiments on the GPU box:
# imd_b200/variants/libimd_b200_NAME.so (git-ignored; travels with gpurun).  Needs `make` first (other objects are shared).
# Use with  IMDB200_LIB=$PWD/imd_b200/variants/libimd_b200_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/var_$name
mkdir -p $out imd_b200/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $*"
$NV -c imd_b200/csrc/forces.cu -o $out/forces.o 2> $out/forces.log &
$NV -DIMDB_CUBIC=1 -c imd_b200/csrc/forces.cu -o $out/forces_cubic.o 2>/dev/null &
$NV -DIMDB_EEAM=1 -c imd_b200/csrc/forces.cu -o $out/forces_eeam.o 2>/dev/null &
$NV -DIMDB_CUBIC=1 -DIMDB_EEAM=1 -c imd_b200/csrc/forces.cu -o $out/forces_cubic_eeam.o 2>/dev/null &
wait
others=$(ls build/*.o | grep -v "build/forces\(_cubic\|_eeam\|_cubic_eeam\)\?\.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o imd_b200/variants/libimd_b200_$name.so $out/*.o $others -lcudart
echo built imd_b200/variants/libimd_b200_$name.so
grep -A2 "k_pass1ILi[0-9]*ELi1ELb1ELb0ELb1ELb0ELb1ELb0ELb0E\|k_pass2ILi[0-9]*ELi1ELb0ELb0ELb1ELb1ELb0ELb0E" $out/forces.log | grep "Used\|spill" | paste - - | sed 's/ptxas info    ://; s/bytes stack frame, //'
