#!/bin/bash
# Experiment helper: build a variant of the library with extra nvcc flags.
#   scripts/build_variant.sh NAME "-DNT_MIN_BLOCKS=3"  ->  nettracer_b200/variants/libnt_NAME.so
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../nettracer_b200/csrc"
OUT=../variants; OBJ=/tmp/nt_var_$NAME
mkdir -p $OUT $OBJ
COMMON="$EXTRA -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -diag-suppress 186"
nvcc $COMMON -fmad=false $PTXAS_V -c nt_kernels_f64.cu -o $OBJ/f64.o &
nvcc $COMMON -use_fast_math $PTXAS_V -c nt_kernels_f32.cu -o $OBJ/f32.o &
nvcc $COMMON -c nt_api.cu -o $OBJ/api.o &
nvcc $COMMON -c nt_peaks.cu -o $OBJ/peaks.o &
nvcc $COMMON -c nt_bvh.cpp -o $OBJ/bvh.o &
nvcc $COMMON -c nt_bvh_gpu.cu -o $OBJ/bvhgpu.o &
nvcc $COMMON -c nt_cull.cpp -o $OBJ/cull.o &
nvcc $COMMON -c nt_shadowgrid.cpp -o $OBJ/shadowgrid.o &
nvcc $COMMON -c nt_multi.cpp -o $OBJ/multi.o &
nvcc $COMMON -c nt_hostframe.cpp -o $OBJ/hostframe.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libnt_$NAME.so $OBJ/f64.o $OBJ/f32.o $OBJ/api.o $OBJ/peaks.o $OBJ/bvh.o $OBJ/bvhgpu.o $OBJ/cull.o $OBJ/shadowgrid.o $OBJ/multi.o $OBJ/hostframe.o -lcudart -lpthread -lrt
echo built $OUT/libnt_$NAME.so
