#!/bin/bash
# Kernel tuning aid (GPU box): correctness (pipelined vs staged gradients) + timing for each lib/variants/<name>.so
#   tools/run_variants.sh name1 name2 ...   -> gpurun_out/variants.log
mkdir -p gpurun_out
for v in "$@"; do
  echo "===== $v" >> gpurun_out/variants.log
  B200INR_LIB=$PWD/mri-super-resolution_b200/lib/variants/$v.so timeout 120 python tools/gpu_dev_check.py ${MODES:-pvs timing} 2>&1 \
    | grep -v "^    seg\|Warning" | awk '/relerr|timing|FAIL|rror|Trace|trap/' >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
