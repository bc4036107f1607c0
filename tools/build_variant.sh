#!/bin/bash
# Builds a variant of libfov360.so with extra nvcc flags (A/B experiments; load with FOV360_LIB=...).
#   tools/build_variant.sh NAME "-DFOO=1 ..."   ->  tools/variants/libfov360_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/variants/obj_$name
C=foveated-360-video_b200/csrc
for f in capi sat_encode sat_onepass sat_decode image_sampler projections color_convert; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math \
       -I include -I $C $@ -c $C/$f.cu -o tools/variants/obj_$name/$f.o &
done
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -I include -I $C -I /usr/local/cuda/include -c $C/luts.cc -o tools/variants/obj_$name/luts.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o tools/variants/libfov360_$name.so tools/variants/obj_$name/*.o
echo tools/variants/libfov360_$name.so
