#!/bin/bash
# One GPU call that produces the round's evidence (run under gpurun from the repo root):
#   tools/profile_round.sh <tag>      -> gpurun_out/<tag>_*.{log,csv,ncu-rep}; copy the summaries into profiles/
# 1. the contract bench line, 2. the ncu launch list of the same command (cold-cache, serialised: compare SHARES),
# 3. one `ncu --set full` capture of the fit step's kernels, 4. the all-config benchmark.
tag=${1:-rN}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.log 2> gpurun_out/${tag}_bench_1gpu.err || exit 1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/gpu_dev_check.py ncu_fit > gpurun_out/${tag}_ncu_fit_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'siren_fwd_kernel|siren_bwdp_kernel' -s 3 -c 3 \
      -f -o gpurun_out/${tag}_fit python tools/gpu_dev_check.py ncu_fit > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/bench_configs.py --steps 10 > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
tail -c 600 gpurun_out/${tag}_bench_1gpu.log
