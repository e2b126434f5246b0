#!/usr/bin/env bash
# developer tool: build an experimental libsss_b200 variant from ONE recompiled source
#   tools/build_variant.sh <name> <source.cu> [-DFLAG=...]...   -> csrc/build/variants/<name>.so
set -e
cd "$(dirname "$0")/../shared_simd_scan_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p build/variants
stem=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC "$@" -c "$src" -o "build/variants/${stem}_$name.o"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "build/variants/$name.so" $(ls build/*.o | grep -v "build/$stem.o") "build/variants/${stem}_$name.o" -ldl -lpthread
echo "built $name"
