#!/bin/bash
# Round-2 GPU session H: launch list of the default bench command; --set full capture of the backward kernels of one
# training step (exported to CSV on the box: the report itself is larger than what gpurun copies back).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/run_train_step.py --steps 2 > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:'attn_bwd|batch_attn_bwd|token_sums|transpose_norm|splitk_reduce|in_bwd_apply' \
    -c 40 -o gpurun_out/r02_bwd python tools/run_train_step.py --steps 1 > gpurun_out/ncu_bwd.log 2>&1
echo "bwd capture exit $?"
ncu -i gpurun_out/r02_bwd.ncu-rep --page raw --csv > gpurun_out/r02_bwd_raw.csv 2> gpurun_out/ncu_export.err
rm -f gpurun_out/r02_bwd.ncu-rep
ls -la gpurun_out; du -sh gpurun_out
