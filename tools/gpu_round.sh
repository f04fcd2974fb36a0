#!/bin/bash
# One GPU-box session: stage tests (one process per group so a faulting kernel cannot hide the others),
# module parity, smoke, stage timings, bench, ncu.  Logs land in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/stages.jsonl
for grp in in_stats linear proj_f32 proj_bf16 attn_f32 attn_bf16; do
  timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "$grp" > gpurun_out/t_$grp.log 2>&1
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
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 6 -c 2 -o gpurun_out/prof_attn \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
  echo "ncu exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/t_*.log
cat gpurun_out/stages.jsonl gpurun_out/bench.json gpurun_out/bench_cfg3.json; tail -5 gpurun_out/bench.err
