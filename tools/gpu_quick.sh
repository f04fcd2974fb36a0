#!/bin/bash
# Short GPU session for kernel iteration: attention stage tests + stage timings.
mkdir -p gpurun_out; rm -f gpurun_out/q_*.log gpurun_out/q_stages.jsonl
timeout 150 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "attn_bf16 or proj_bf16 or linear" > gpurun_out/q_attn.log 2>&1; echo "attn tests exit $?"
for cfg in "--B 8 --hw 64" "--B 1 --hw 128" "--B 1 --hw 64"; do
  timeout 300 python tools/bench_stages.py $cfg >> gpurun_out/q_stages.jsonl 2>> gpurun_out/q_stages.err
done
tail -n 3 gpurun_out/q_attn.log; cat gpurun_out/q_stages.jsonl
if [ "$1" == "full" ]; then
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -m gpu -q -x > gpurun_out/q_all.log 2>&1; echo "all tests exit $?"; tail -n 8 gpurun_out/q_all.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/q_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/q_smoke.log
fi
