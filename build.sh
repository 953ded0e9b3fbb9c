#!/usr/bin/env bash
# Build quantum_b200/libtfqb.so (C ABI + sm_100a kernels) in-tree.
set -euo pipefail
cd "$(dirname "$0")/quantum_b200/csrc"
OUT=../libtfqb.so
OBJ=${TFQB_OBJ_DIR:-../../build/obj}
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXXFLAGS="-std=c++17 -O3 -fPIC -Wall"
NVFLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
# text of the device helpers, embedded for the run-time specialised kernels
{ printf 'R"JITSRC('; cat pass_device.cuh; printf ')JITSRC"\n'; } > "$OBJ/pass_device_src.inc"
pids=()
for f in wire program plan ps_ops; do
  g++ $CXXFLAGS -c $f.cc -o "$OBJ/$f.o" & pids+=($!)
done
g++ $CXXFLAGS -I/usr/local/cuda/include -I"$OBJ" -c jit.cc -o "$OBJ/jit.o" & pids+=($!)
$NVCC $NVFLAGS -c kernels.cu -o "$OBJ/kernels.o" & pids+=($!)
$NVCC $NVFLAGS -c backend.cu -o "$OBJ/backend.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" \
  "$OBJ/wire.o" "$OBJ/program.o" "$OBJ/plan.o" "$OBJ/ps_ops.o" "$OBJ/jit.o" "$OBJ/kernels.o" "$OBJ/backend.o" -ldl
echo "built $(readlink -f $OUT)"
