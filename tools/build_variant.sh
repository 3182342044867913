#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh NAME [-DMACRO=VALUE ...]  -> build/variants/libpa_NAME.so
NAME=$1; shift
mkdir -p build/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-pthread -shared "$@" \
     -o build/variants/libpa_$NAME.so pyascore_b200/csrc/pa_lib.cu pyascore_b200/csrc/pa_host.cpp && echo built $NAME
