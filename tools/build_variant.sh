#!/bin/bash
# tools/build_variant.sh name [-Dflags...]  ->  build_variants/lib_<name>.so
name=$1; shift
mkdir -p build_variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -Wno-deprecated-declarations -shared "$@" dexdeform_b200/csrc/abi1_kernels.cu dexdeform_b200/csrc/engine.cu dexdeform_b200/csrc/fk.cu -o build_variants/lib_$name.so
