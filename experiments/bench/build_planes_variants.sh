#!/bin/bash
# Build libgppvae_b200 variants of the operand-plane kernels (accumulation window length) into experiments/bench/variants/.
# usage: build_planes_variants.sh 16 32 64
set -e
cd "$(dirname "$0")/../.."
mkdir -p experiments/bench/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr"
python -m gppvae_b200.build > /dev/null
for v in "$@"; do
  nvcc $FLAGS $EXTRA_DEFS -DGPP_PL_WIN=$v -c gppvae_b200/csrc/gemm_planes.cu -o /tmp/gemm_planes_w$v.o
  objs=$(ls gppvae_b200/build/*.o | grep -v "gemm_planes.o")
  nvcc -shared -o experiments/bench/variants/lib_plwin$v.so $objs /tmp/gemm_planes_w$v.o -gencode arch=compute_100a,code=sm_100a -cudart shared
  echo built plwin$v
done
