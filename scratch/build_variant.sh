#!/bin/bash
# Builds a variant of the library that differs in ONE translation unit's -D flags (kernel-tuning experiments):
#   scratch/build_variant.sh NAME msda_launch_win "-DMSDA_WIN_CELL_COST=0"
# -> richsem_b200/lib/variants/libmsda_NAME.so (other objects are taken from the regular build); use with MSDA_B200_LIB.
set -e
name=$1; unit=$2; flags=$3
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/richsem_b200/lib/obj; out=$root/richsem_b200/lib/variants; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -c -o $out/${unit}_$name.o $root/richsem_b200/csrc/$unit.cu
others=$(ls $obj/*.o | grep -v "/$unit.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmsda_$name.so $out/${unit}_$name.o $others
rm -f $out/${unit}_$name.o
echo $out/libmsda_$name.so
