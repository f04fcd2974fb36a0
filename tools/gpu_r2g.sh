#!/bin/bash
# Round-2 GPU session G: the driver's own checks (pytest -m gpu -x, smoke, default bench), launch list of the default
# bench command, --set full capture of the backward kernels of one training step.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err; echo "cfg5 exit $?" >> gpurun_out/summary.txt
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
python tools/run_train_step.py --steps 2 > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd|batch_attn_bwd|token_sums|transpose_norm|splitk_reduce|in_bwd_apply' \
    -c 60 -o gpurun_out/r02_bwd python tools/run_train_step.py --steps 1 > gpurun_out/ncu_bwd.log 2>&1
echo "bwd capture exit $?" >> gpurun_out/summary.txt
ncu -i gpurun_out/r02_bwd.ncu-rep --page raw --csv > gpurun_out/r02_bwd_raw.csv 2> gpurun_out/ncu_export.err
ls -la gpurun_out/r02_bwd.ncu-rep
cat gpurun_out/summary.txt
grep -E "passed|failed|error" gpurun_out/t_all.log | tail -3; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/t_all.log | head -20
tail -2 gpurun_out/smoke.log; cut -c1-1500 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench.err
