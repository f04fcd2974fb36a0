#!/usr/bin/env python
"""Development aid: where the three roles of gemm_tc_kernel wait (needs a build with MHADA_NVCC_EXTRA=-DMHADA_GEMM_TRACE).
Prints, averaged over the CTAs: MMA-issuer loop cycles and the part spent waiting for operands (full) / for the epilogue
(acc_empty); producer cycles waiting for free stages; epilogue cycles waiting for accumulators / for store slabs."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mhada_style_transfer_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
M = 32768
for name, N, K, mode in (("in_proj", 1536, 512, "bf16"), ("fc1", 2048, 512, "bf16"), ("out_conv", 512, 512, "bf16"),
                         ("fc2", 512, 2048, "bf16")):
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16(); b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.mhada_gemm_bf16(P(x), K, P(w), K, P(b), M, N, K, P(y), N, None, N, None, 0, 0, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.mhada_gemm_bf16(P(x), K, P(w), K, P(b), M, N, K, P(y), N, None, N, None, 0, 0, 0, st); e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (148 * 8))()
    rc = L.mhada_debug_gemm_trace(buf, 148)
    a = np.array(buf[:], dtype=np.float64).reshape(148, 8)
    lead = a[a[:, 0] > 0]                 # CTAs that issued MMAs (leaders in pair mode)
    print(json.dumps({"gemm": name, "us": round(e0.elapsed_time(e1) * 1e3, 1), "issuers": int(len(lead)),
                      "items_per_issuer": round(float(lead[:, 7].mean()), 2),
                      "mma_loop_cyc": int(lead[:, 0].mean()), "mma_wait_full": int(lead[:, 2].mean()),
                      "mma_wait_acc_empty": int(lead[:, 1].mean()), "producer_wait_empty": int(a[:, 3].mean()),
                      "epi_loop_cyc": int(a[:, 4].mean()), "epi_wait_acc_full": int(a[:, 5].mean()),
                      "epi_wait_store_slab": int(a[:, 6].mean())}))
