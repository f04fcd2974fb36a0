#!/bin/bash
# Round-2 GPU session D: full tests, ViT timings, bench (default + cfg5 + cfg4 + cfg1).
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_vit test_gpu_stages test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_vit.py --B 8 --img 512 > gpurun_out/vit_b8.json 2> gpurun_out/vit.err
timeout 300 python tools/bench_vit.py --B 1 --img 1024 > gpurun_out/vit_b1.json 2>> gpurun_out/vit.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.log; do echo "== $f"; grep -E "passed|failed|error" $f | tail -3; grep -E "^(FAILED|ERROR)" $f | head -20; done
tail -3 gpurun_out/smoke.log; cat gpurun_out/vit_b8.json gpurun_out/vit_b1.json; tail -5 gpurun_out/vit.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
cat gpurun_out/bench_cfg5.json; tail -5 gpurun_out/bench_cfg5.err; cat gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err; cat gpurun_out/bench_cfg1.json; tail -3 gpurun_out/bench_cfg1.err
