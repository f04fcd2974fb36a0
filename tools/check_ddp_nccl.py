"""Data-parallel training step over NCCL (BASELINE configs[4]: gradient all-reduce over NVLink): every rank runs
forward + backward on its shard of the batch, `sharding.allreduce_gradients` averages the gradients, rank 0 compares
with the full-batch gradients.  Development aid / multi-GPU check:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_nccl.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhada_style_transfer_b200 as M  # noqa: E402
from mhada_style_transfer_b200.network import set_precision  # noqa: E402
from mhada_style_transfer_b200.sharding import allreduce_gradients, shard_range  # noqa: E402
from oracle import cases, synth  # noqa: E402  (seeded inputs only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 2 * world
    case = dict(B=B, hw=(8, 8), hsws=(8, 8), seed=96)
    fc, fs, sd = cases.transformer_inputs(case)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)

    def build():
        m = M.AdaAttnTransformerMultiHead()
        m.load_state_dict(synth.to_torch(sd, torch.float32), strict=True)
        return set_precision(m.to(dev).train(), "bf16")

    m = build()
    s, e = shard_range(B, rank, world)
    _, cs = m([to(x[s:e]) for x in fc], [to(x[s:e]) for x in fs])
    (cs.float().sum() / B).backward()
    for p in m.parameters():
        p.grad.mul_(world)                      # sum of the shard gradients = full-batch gradient
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    allreduce_gradients(m)
    t1.record()
    torch.cuda.synchronize()
    if rank == 0:
        full = build()
        _, cs2 = full([to(x) for x in fc], [to(x) for x in fs])
        (cs2.float().sum() / B).backward()
        scale = max(b.grad.abs().max().item() for b in full.parameters())
        worst = max((a.grad - b.grad).abs().max().item() / (b.grad.abs().max().item() + 1e-3 * scale)
                    for a, b in zip(m.parameters(), full.parameters()))
        nbytes = sum(p.grad.numel() * 4 for p in m.parameters())
        print(f"ranks {world}: worst relative gradient difference {worst:.2e}; all-reduce of {nbytes / 1e6:.1f} MB in "
              f"{t0.elapsed_time(t1):.2f} ms", flush=True)
        assert worst <= 5e-3, worst
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
