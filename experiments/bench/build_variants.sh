#!/bin/bash
# Build libgppvae_b200 variants of the tcgen05 kernels (window / group / stage counts) into experiments/bench/variants/.
# usage: build_variants.sh "W G RAW [prof|-] [LO]" ...
set -e
cd "$(dirname "$0")/../.."
mkdir -p experiments/bench/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr"
python -m gppvae_b200.build > /dev/null
for v in "$@"; do
  set -- $v
  name="w$1g$2r$3l${5:-4}"; if [ "$4" = "prof" ]; then name="${name}prof"; fi
  extra=""
  if [ "$4" = "prof" ]; then extra="-DGPP_TC_PROF"; fi
  nvcc $FLAGS $extra $EXTRA_DEFS -DGPP_TC_WINGROUPS=$1 -DGPP_TC_GROUP=$2 -DGPP_TC_RAW=$3 -DGPP_TC_LO=${5:-4} -c gppvae_b200/csrc/gemm_tc.cu -o /tmp/gemm_tc_$name.o
  objs=$(ls gppvae_b200/build/*.o | grep -v "gemm_tc.o")
  nvcc -shared -o experiments/bench/variants/lib_$name.so $objs /tmp/gemm_tc_$name.o -gencode arch=compute_100a,code=sm_100a -cudart shared
  echo built $name
done
