#!/bin/bash
# tools/build_variant.sh NAME SOURCE.cu [extra nvcc flags...]: variants/NAME.so = the library with csrc/SOURCE.cu recompiled
# under the extra flags (A/B runs on one GPU box: CZK_B200_LIB=variants/NAME.so python tools/msm_once.py ...)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift; shift
P=collaborative-zksnark_b200
mkdir -p variants
python $P/build.py > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden \
  -I include --expt-relaxed-constexpr "$@" -c $P/csrc/$src -o variants/$name.o
objs=$(ls $P/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o variants/$name.so $objs variants/$name.o -cudart static -ldl -lpthread -lrt
echo variants/$name.so
