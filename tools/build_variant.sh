#!/bin/bash
# Development: build a variant of libzstdlite_gpu.so with extra -D switches into variants/<name>.so (A/B runs on one box:
# ZSTDLITE_GPU_LIB=variants/<name>.so python bench.py ...).  usage: tools/build_variant.sh name -DZL_EXEC_V1=1 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/zstdlite_b200/csrc
tmp=$(mktemp -d)
mkdir -p $root/variants
for f in zl_api zl_dec_kernels zl_api_compress zl_enc_kernels; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden "$@" -I$src -c $src/$f.cu -o $tmp/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/variants/$name.so $tmp/*.o -Xlinker -Bsymbolic
rm -rf $tmp
echo built variants/$name.so
