#!/bin/bash
# Builds a tuning variant of libminotert.so with extra nvcc flags into variants/<name>/ (git-ignored;
# travels to the GPU box).  Usage: tools/build_variant.sh <name> "<extra nvcc flags>"
# Select it at run time with MINOTERT_LIB_DIR=variants/<name>.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXTRA=$2
OUT=$ROOT/variants/$NAME
mkdir -p "$OUT/obj"
cd "$ROOT/minotert_b200/csrc"
for f in api sky spheres tonemap denoise temporal sort bvh_build mesh group probe; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
    -Xcompiler -fPIC -ccbin /usr/bin/g++ $EXTRA -c $f.cu -o "$OUT/obj/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libminotert.so" "$OUT"/obj/*.o -ccbin /usr/bin/g++ -ldl
cp "$ROOT/minotert_b200/libminote_host.so" "$OUT/"
rm -rf "$OUT/obj"
echo "built $OUT"
