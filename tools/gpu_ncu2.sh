#!/bin/bash
mkdir -p gpurun_out
python tools/bench_stages.py --B 8 --hw 64 --iters 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"proj_tc|linear_tc|stats_partial|fold_stats" -s 4 -c 8 -o gpurun_out/prof_small -f \
    python tools/bench_stages.py --B 8 --hw 64 --iters 2 > gpurun_out/ncu3.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu3.log
