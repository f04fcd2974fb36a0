#!/bin/bash
# Round-2 GPU session E: halo conv (both descriptor interpretations), GEMM TMA epilogue, timings, bench.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
MHADA_CONV_HALO=1 timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "conv3x3_tc" > gpurun_out/t_conv_halo1.log 2>&1; echo "conv halo=1 exit $?" >> gpurun_out/summary.txt
MHADA_CONV_HALO=2 timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "conv3x3_tc" > gpurun_out/t_conv_halo2.log 2>&1; echo "conv halo=2 exit $?" >> gpurun_out/summary.txt
MHADA_CONV_HALO=0 timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "conv3x3_tc" > gpurun_out/t_conv_halo0.log 2>&1; echo "conv halo=0 exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_vit.py -m gpu -q > gpurun_out/t_vit.log 2>&1; echo "vit exit $?" >> gpurun_out/summary.txt
H=1; grep -q "failed" gpurun_out/t_conv_halo1.log && H=2; grep -q "failed" gpurun_out/t_conv_halo${H}.log && H=0
echo "using MHADA_CONV_HALO=$H" >> gpurun_out/summary.txt
MHADA_CONV_HALO=$H timeout 300 python tools/bench_decoder.py 8 64 > gpurun_out/decoder_b8.json 2> gpurun_out/decoder.err
MHADA_CONV_HALO=0 timeout 300 python tools/bench_decoder.py 8 64 > gpurun_out/decoder_b8_nohalo.json 2>> gpurun_out/decoder.err
timeout 300 python tools/bench_vit.py --B 8 --img 512 > gpurun_out/vit_b8.json 2> gpurun_out/vit.err
MHADA_CONV_HALO=$H timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/t_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/summary.txt
MHADA_CONV_HALO=$H timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.log; do echo "== $f"; grep -E "passed|failed|error" $f | tail -3; grep -E "^(FAILED|ERROR)" $f | head -20; done
cat gpurun_out/decoder_b8.json gpurun_out/decoder_b8_nohalo.json; tail -3 gpurun_out/decoder.err; cat gpurun_out/vit_b8.json; tail -3 gpurun_out/vit.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
