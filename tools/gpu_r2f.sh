#!/bin/bash
# Round-2 GPU session F: forloss tensor-core path, full tests, cfg5 sweep, bench.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "forloss" > gpurun_out/t_forloss.log 2>&1; echo "forloss exit $?" >> gpurun_out/summary.txt
for f in test_gpu_vit test_gpu_stages test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/sweep_cfg5.py > gpurun_out/sweep_cfg5.jsonl 2> gpurun_out/sweep_cfg5.err; echo "sweep exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.log; do echo "== $f"; grep -E "passed|failed|error" $f | tail -3; grep -E "^(FAILED|ERROR)|^E  " $f | head -20; done
tail -3 gpurun_out/smoke.log; cat gpurun_out/sweep_cfg5.jsonl; tail -3 gpurun_out/sweep_cfg5.err; cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
