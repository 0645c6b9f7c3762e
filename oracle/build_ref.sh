#!/usr/bin/env bash
# Builds oracle/_ref/libdogm_ref.so: the reference's own CUDA implementation of the DOGM cycle, compiled from the
# sources where they lie under /root/reference (never copied), plus oracle/ref_harness.cu.  TEST INFRASTRUCTURE.
# The reference's CMake cannot be used here (needs OpenGL/GLFW/GLEW/GLM/OpenCV and downloads googletest at configure
# time, dogm/CMakeLists.txt:9-31); these are the ten .cu files of the `dogm` library target plus the measurement-grid
# kernels, built with the flags the reference's CMake would pass (C++14, --expt-extended-lambda) for sm_100.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${DOGM_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/dogm/src" ]; then
    echo "reference sources not found under $REF (the GPU box only uses the prebuilt $OUT/libdogm_ref.so)"
    exit 3
fi
mkdir -p "$OUT/obj"
NVCC="${NVCC:-nvcc}"
FLAGS=(-std=c++14 -O2 --expt-extended-lambda -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -w
       -I"$HERE/glm_shim" -I"$REF/dogm/include" -I"$REF/dogm/demo/simulator/include")
SRCS=("$REF"/dogm/src/dogm.cu "$REF"/dogm/src/kernel/*.cu "$REF"/dogm/demo/simulator/mapping/kernel/measurement_grid.cu
      "$HERE"/ref_harness.cu)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
    o="$OUT/obj/$(basename "${s%.cu}").o"
    OBJS+=("$o")
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100,code=sm_100 -o "$OUT/libdogm_ref.so" "${OBJS[@]}"
rm -rf "$OUT/obj"
echo "built $OUT/libdogm_ref.so"
