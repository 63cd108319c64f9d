#!/bin/bash
# Kernel tuning aid: build libb200inr with extra -D flags for ONE translation unit into lib/variants/<name>.so
#   tools/build_variant.sh <name> <unit.cu> -DFOO=1 ...      then run with B200INR_LIB=<path>
set -e
name=$1; unit=$2; shift 2
cd "$(dirname "$0")/../mri-super-resolution_b200/csrc"
make -s >/dev/null
mkdir -p build/var ../lib/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v \
  --expt-relaxed-constexpr "$@" -c "$unit" -o "build/var/${name}.o" 2> "build/var/${name}.log"
grep -A2 "siren_bwdp_kernelILb0\|siren_fwd_kernelILi256ELi2" "build/var/${name}.log" | grep "Used\|spill" || true
objs=$(ls build/*.o | grep -v "build/${unit%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "../lib/variants/${name}.so" $objs "build/var/${name}.o" -lcudart
echo "built lib/variants/${name}.so"
