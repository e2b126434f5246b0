#!/usr/bin/env bash
# build_ref.sh -- compile the UNMODIFIED reference hot path, once per bit width, into
# oracle/_ref/libref_wNN.so (git-ignored; travels to the GPU box with the gpurun snapshot).
#
# TEST INFRASTRUCTURE ONLY.  The reference's sources are read where they lie under
# $REF (/root/reference/src); nothing is copied into the repository.  The reference fixes its bit
# width with `#define BITS_NEEDED 9` (src/simd_scan.hpp:12, no #ifndef guard, so -D cannot
# override it); for the other widths the header is piped through sed into a throw-away temp dir
# that is searched first, and each .cpp is fed to the compiler on stdin so that its
# `#include "simd_scan.hpp"` resolves to that substituted header instead of its sibling file.
# The reference's cmake build is not used (it needs workarounds on this image and builds the
# benchmark CLI we do not want); this recipe is the whole build.
#
# usage: oracle/build_ref.sh [width ...]      (default: 1..32)
set -euo pipefail

REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
CXX=${REF_CXX:-/usr/bin/g++}     # /opt/gcc/bin/g++ cannot link -fopenmp on this image
FLAGS="-std=gnu++17 -O3 -mssse3 -msse4.1 -mavx -mavx2 -fopenmp -fPIC -w"
UNITS="simd_scan simd_scan_compression simd_scan_decompression simd_scan_shared simd_scan_shared_linear util"

if [ ! -d "$REF/src" ]; then
    echo "build_ref: $REF/src not present (GPU box?) -- keeping prebuilt oracle/_ref as is"
    exit 0
fi
mkdir -p "$OUT"

build_width() {
    local w=$1
    local so="$OUT/libref_w$(printf %02d "$w").so"
    if [ -f "$so" ] && [ "$so" -nt "$HERE/ref_shim.cpp" ] && [ "$so" -nt "$HERE/build_ref.sh" ]; then return 0; fi
    local tmp; tmp="$(mktemp -d)"
    sed "s/#define BITS_NEEDED 9/#define BITS_NEEDED $w/" "$REF/src/simd_scan.hpp" > "$tmp/simd_scan.hpp"
    (
        cd "$tmp"
        for u in $UNITS; do
            $CXX $FLAGS -x c++ -iquote "$tmp" -iquote "$REF/src" -c - -o "$tmp/$u.o" < "$REF/src/$u.cpp"
        done
        $CXX $FLAGS -iquote "$tmp" -iquote "$REF/src" -c "$HERE/ref_shim.cpp" -o "$tmp/ref_shim.o"
        $CXX -shared -fopenmp -o "$so.tmp" "$tmp"/*.o -lpthread
        mv "$so.tmp" "$so"
    )
    rm -rf "$tmp"
    echo "built $so"
}

WIDTHS=("$@")
if [ ${#WIDTHS[@]} -eq 0 ]; then WIDTHS=($(seq 1 32)); fi

JOBS=${JOBS:-$(nproc)}
running=0
for w in "${WIDTHS[@]}"; do
    build_width "$w" &
    running=$((running + 1))
    if [ "$running" -ge "$JOBS" ]; then wait -n; running=$((running - 1)); fi
done
wait
