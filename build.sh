#!/bin/bash
# Build libsidekit_b200.so (sm_100a) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=sidekit_b200/csrc
OUT=sidekit_b200/libsidekit_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$SRC"
mkdir -p build
objs=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ -n "$(find $SRC include -name '*.cuh' -newer $o -o -name '*.h' -newer $o)" ]; then
    $NVCC $FLAGS "$@" -c $f -o $o &
  fi
  objs="$objs $o"
done
wait
$NVCC -shared -o $OUT $objs -lcudart
echo "built $OUT"
