#!/bin/bash
# Build and time attention-kernel variants on the GPU box (nvcc is in the image).
#   tools/gpu_variants.sh "<nvcc -D flags of variant 1>" "<variant 2>" ...
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  MHADA_NVCC_EXTRA="$v" python -m mhada_style_transfer_b200.build --force > gpurun_out/v_build_$i.log 2>&1 || { echo "build $v failed"; tail -5 gpurun_out/v_build_$i.log; continue; }
  echo "== variant $v"
  timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "attn_bf16" 2>&1 | tail -1
  timeout 300 python tools/bench_stages.py --B 8 --hw 64 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('attn_ms','attn_tflops','layer_ms')})"
  timeout 300 python tools/bench_stages.py --B 1 --hw 128 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('attn_ms','attn_tflops','layer_ms')})"
  if [ -n "$TRACE" ]; then timeout 120 python tools/trace_attn.py 2>&1 | tail -22; fi
done
