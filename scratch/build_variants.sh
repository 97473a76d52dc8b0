#!/bin/bash
# builds variants of the library that differ in the tile kernel's ring geometry (3D order 3 translation unit only):
#   scratch/variants/libzfvm_<name>.so ; usage: build_variants.sh name "-DZFVM_SLOT_TARGET=3072 -DZFVM_NSLOTS=5" [unit]
set -e
cd "$(dirname "$0")/../zisafvm_b200/csrc"
name=$1; flags=$2; unit=${3:-recon_3d_deg2}
mkdir -p ../../scratch/variants
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fopenmp,-O3 -Xptxas -v $flags \
  -c kernels/$unit.cu -o ../../scratch/variants/${unit}_$name.o 2> ../../scratch/variants/${unit}_$name.ptxas.log
objs=$(ls build/kernels/*.o build/*.o build/host/*.o | grep -v "kernels/$unit.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o ../../scratch/variants/libzfvm_$name.so $objs ../../scratch/variants/${unit}_$name.o \
  /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a -lcudart -lgomp -ldl
grep -E "spill|registers" ../../scratch/variants/${unit}_$name.ptxas.log | sort | uniq -c | sort -rn | head -4
