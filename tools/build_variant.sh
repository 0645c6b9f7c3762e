#!/bin/bash
# Builds an A/B variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]  ->  build_ab/lib_<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
SRC=$ROOT/dynamic-occupancy-grid-map_b200/csrc
OUT=$ROOT/build_ab
mkdir -p $OUT/obj_$NAME
FLAGS="-O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static $*"
for f in dogm_api kernels_particles kernels_cells kernels_meas band_group; do
  nvcc $FLAGS -c $SRC/$f.cu -o $OUT/obj_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $OUT/lib_$NAME.so $OUT/obj_$NAME/*.o
rm -rf $OUT/obj_$NAME
echo $OUT/lib_$NAME.so
