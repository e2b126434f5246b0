#!/usr/bin/env bash
# developer tool: rebuild EVERY CUDA source with extra flags into csrc/build/variants/<name>.so
#   tools/build_variant_all.sh <name> [-DFLAG=...]...
set -e
cd "$(dirname "$0")/../shared_simd_scan_b200/csrc"
name=$1; shift
mkdir -p build/variants/$name
for f in *.cu; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC "$@" -c "$f" -o "build/variants/$name/${f%.cu}.o" ) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "build/variants/$name.so" build/variants/$name/*.o -ldl -lpthread
rm -rf "build/variants/$name"
echo "built $name"
