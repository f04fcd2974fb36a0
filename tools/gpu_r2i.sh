#!/bin/bash
# Round-2 GPU session I (final state): the driver's checks, default bench, cfg5, launch list, decoder / ViT stage timings.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2>> gpurun_out/bench.err; echo "cfg5 exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --workload cfg4 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2>> gpurun_out/bench.err; echo "cfg4 exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/bench_decoder.py > gpurun_out/decoder_blocks_b8.json 2>> gpurun_out/bench.err
timeout 200 python tools/bench_vit.py > gpurun_out/vit_b8_stage_times.json 2>> gpurun_out/bench.err
timeout 200 python tools/bench_latency.py > gpurun_out/latency.json 2>> gpurun_out/bench.err
timeout 300 python tools/sweep_cfg5.py > gpurun_out/sweep_cfg5.jsonl 2>> gpurun_out/bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-also > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|error" gpurun_out/t_all.log | tail -3; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/t_all.log | head -20
tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench.err
du -sh gpurun_out
