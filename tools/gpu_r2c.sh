#!/bin/bash
# Round-2 GPU session C: conv kernel tests, decoder timings, full tests, bench.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "conv3x3_tc" > gpurun_out/t_conv.log 2>&1; echo "conv exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_decoder.py 8 64 > gpurun_out/decoder_b8.json 2> gpurun_out/decoder.err
timeout 300 python tools/bench_decoder.py 1 128 > gpurun_out/decoder_b1.json 2>> gpurun_out/decoder.err
for f in test_gpu_vit test_gpu_stages test_gpu_parity; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/sanitize_cases.py > gpurun_out/edge_cases.log 2>&1; echo "edge cases exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.log; do echo "== $f"; grep -E "passed|failed|error" $f | tail -3; grep -E "^(FAILED|ERROR)" $f | head -20; done
tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/edge_cases.log; cat gpurun_out/decoder_b8.json gpurun_out/decoder_b1.json; tail -5 gpurun_out/decoder.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
