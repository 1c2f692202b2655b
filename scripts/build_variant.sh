#!/bin/bash
# A/B builds of the fused step kernel: scripts/build_variant.sh <name> [extra nvcc flags for ble_step_fused.cu]
# -> balloon_learning_environment_b200/variants/libble_<name>.so (use with BLE_B200_LIB=<path>); git-ignored.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
pkg=balloon_learning_environment_b200
mkdir -p $pkg/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o $pkg/variants/fused_$name.o $pkg/csrc/ble_step_fused.cu
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o $pkg/variants/libble_$name.so $pkg/build/ble_engine.o $pkg/variants/fused_$name.o $pkg/build/ble_learner.o $pkg/build/ble_dense.o -lcublasLt -Xlinker -rpath=/usr/local/cuda/lib64
rm -f $pkg/variants/fused_$name.o
echo $pkg/variants/libble_$name.so
