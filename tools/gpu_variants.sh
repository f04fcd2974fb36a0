#!/bin/bash
# Build and time attention-kernel variants on the GPU box (nvcc is in the image).
mkdir -p gpurun_out
i=0
for v in "-DMHADA_AT_PINGPONG=2" "-DMHADA_AT_PINGPONG=0" "-DMHADA_AT_PINGPONG=2 -DMHADA_AT_EPI_ROLLED" "-DMHADA_AT_PINGPONG=2 -DMHADA_AT_FORCE_PERSISTENT=0" "-DMHADA_AT_PINGPONG=0 -DMHADA_AT_FORCE_PERSISTENT=0 -DMHADA_AT_EPI_ROLLED"; do
  i=$((i+1))
  MHADA_NVCC_EXTRA="$v" python -m mhada_style_transfer_b200.build --force > gpurun_out/v_build_$i.log 2>&1 || { echo "build $v failed"; tail -5 gpurun_out/v_build_$i.log; continue; }
  echo "== variant $v"
  timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "attn_bf16" 2>&1 | tail -1
  timeout 300 python tools/bench_stages.py --B 8 --hw 64 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('attn_ms','attn_tflops','layer_ms')})"
  timeout 300 python tools/bench_stages.py --B 1 --hw 128 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('attn_ms','attn_tflops','layer_ms')})"
done
