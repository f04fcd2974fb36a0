#!/bin/bash
# ncu --set full of the HBM-bound stage kernels (one launch each, from the stand-alone stage bench)
mkdir -p gpurun_out
python tools/bench_stages.py --B 8 --hw 64 --iters 2 > gpurun_out/plain2.log 2>&1 || exit 1
for k in proj_tc_kernel linear_tc_kernel fold_stats_kernel stats_partial_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_$k -f \
      python tools/bench_stages.py --B 8 --hw 64 --iters 2 > gpurun_out/ncu3_$k.log 2>&1
  echo "ncu $k exit $?"
done
