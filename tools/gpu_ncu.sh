#!/bin/bash
# ncu captures for profiles/: launch list of one bench run + full-set capture of the attention kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "attn_bf16" > gpurun_out/n_attn.log 2>&1; echo "attn tests exit $?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 6 -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu attn exit $?"
tail -3 gpurun_out/ncu2.log
