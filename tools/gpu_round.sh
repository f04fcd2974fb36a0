#!/bin/bash
# One GPU-box session: stage tests (one process per group so a faulting kernel cannot hide the others),
# module parity, smoke, stage timings.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for grp in in_stats linear proj_f32 proj_bf16 attn_f32 attn_bf16; do
  timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "$grp" > gpurun_out/t_$grp.log 2>&1
  echo "$grp exit $?" >> gpurun_out/summary.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
for cfg in "--B 8 --hw 64" "--B 1 --hw 128" "--B 1 --hw 64"; do
  timeout 300 python tools/bench_stages.py $cfg >> gpurun_out/stages.jsonl 2>> gpurun_out/stages.err
done
timeout 300 python tools/bench_stages.py --B 1 --hw 64 --dtype fp32 --iters 3 >> gpurun_out/stages.jsonl 2>> gpurun_out/stages.err
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/t_*.log
cat gpurun_out/stages.jsonl
