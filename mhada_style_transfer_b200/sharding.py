"""Image / frame sharding across the GPUs of one box.

The MHAda path has no exchange step: every image (or video frame, given the style features) is independent
(`AdaAttnMultiHead.forward` is batch-independent, adaDecoder.py:162-206; infer_video.py:73-118 processes frames
one by one).  So the batch is partitioned contiguously over ranks, one process per GPU, and the ONLY collective
is the gather of the results.  Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU
tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced partition of range(n_items): the first n % world ranks get one extra item."""
    if n_items < 0 or world_size <= 0:
        raise ValueError("n_items must be >= 0 and world_size > 0")
    base, extra = divmod(n_items, world_size)
    bounds, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        bounds.append((start, start + size))
        start += size
    return bounds


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    return shard_bounds(n_items, world_size)[rank]


def gather_batches(local: torch.Tensor, n_items: int, dst: int = 0, group=None):
    """Gather per-rank result shards (dim 0 = images of this rank's shard_range) into the full batch on `dst`.
    Shards may be uneven or empty: they are padded to the largest shard for the collective and trimmed after.
    Returns the (n_items, ...) tensor on rank `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = shard_bounds(n_items, world)
    largest = max(e - s for s, e in bounds)
    if largest == 0:
        return local[:0] if rank == dst else None
    pad = largest - local.shape[0]
    send = local.contiguous()
    if pad:
        send = torch.cat([send, send.new_zeros((pad,) + tuple(send.shape[1:]))], dim=0)
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[: e - s] for b, (s, e) in zip(bufs, bounds)], dim=0)


def run_sharded(model: Callable, fc: Sequence[torch.Tensor], fs: Sequence[torch.Tensor], dst: int = 0, group=None):
    """`model(fc, fs) -> (fcs, cs)` on this rank's slice of the batch; decoded images gathered on `dst`.
    fc / fs are lists of (B, C, h, w) feature maps holding the WHOLE batch (every rank passes the same B)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = fc[0].shape[0]
    s, e = shard_range(n, rank, world)
    if e > s:
        _, cs = model([t[s:e] for t in fc], [t[s:e] for t in fs])
    else:  # more ranks than images: contribute an empty shard of the right trailing shape
        with torch.no_grad():
            _, probe = model([t[:1] for t in fc], [t[:1] for t in fs])
        cs = probe[:0]
    return gather_batches(cs, n, dst=dst, group=group)


def allreduce_gradients(module: torch.nn.Module, group=None, bucket_bytes: int = 32 << 20) -> None:
    """Average the gradients of `module` over the ranks (the data-parallel step train_image.py never had: it is
    single-GPU, train_image.py:139-144).  Gradients are flattened into ~32 MB buckets so the 25 M parameters of the
    three networks take a handful of collectives; over NVLink 5 the bucket size is chosen for launch latency and
    overlap, not for link count."""
    world = dist.get_world_size(group)
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


class OverlappedGradientAllReduce:
    """Gradient averaging that OVERLAPS the backward pass (BASELINE configs[4]: train step with gradient all-reduce
    over NVLink): parameters are grouped into ~`bucket_bytes` buckets in reverse registration order (the order in
    which backward produces gradients); a post-accumulate-grad hook per parameter marks it ready, and as soon as a
    bucket is complete its flattened gradients go out as one asynchronous all-reduce (NCCL runs it on its own
    stream while autograd keeps computing).  `finish()` -- called between backward() and optimizer.step() -- waits
    for the collectives and scatters the averages back into the .grad tensors.

        sync = OverlappedGradientAllReduce([vit_c, vit_s, adaFormer])
        loss.backward(); sync.finish(); opt.step()

    The reference trains on one GPU (train_image.py:139-144: backward, three Adam steps, no synchronisation)."""

    def __init__(self, modules, group=None, bucket_bytes: int = 32 << 20):
        if isinstance(modules, torch.nn.Module):
            modules = [modules]
        self.group = group
        self.world = dist.get_world_size(group)
        params = [p for m in modules for p in m.parameters() if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._inflight = []            # (bucket index, flat tensor, work handle)
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.launched_in_backward = 0  # buckets whose all-reduce started before finish() (i.e. overlapped)
        self.enabled = True            # False: hooks and finish() do nothing (timing a step without its collectives)

    def _launch(self, i):
        grads = [p.grad for p in self.buckets[i] if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((i, flat, work))

    def _on_grad(self, p):
        if not self.enabled:
            return
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)
            self.launched_in_backward += 1

    def finish(self):
        """Wait for the collectives, write the averaged gradients back, re-arm for the next step."""
        if not self.enabled:
            return
        for i, n in enumerate(self._pending):          # buckets with parameters that got no gradient this step
            if 0 < n < len(self.buckets[i]) or (n == len(self.buckets[i]) and any(p.grad is not None for p in self.buckets[i])):
                self._launch(i)
        for i, flat, work in self._inflight:
            work.wait()
            flat.div_(self.world)
            grads, views, off = [], [], 0
            for p in self.buckets[i]:
                if p.grad is None:
                    continue
                n = p.grad.numel()
                grads.append(p.grad)
                views.append(flat[off: off + n].view_as(p.grad))
                off += n
            # ONE multi-tensor copy per bucket: a copy_ per parameter is ~400 launches at the very end of the step,
            # where nothing is left to hide them behind (1.4 of the 1.6 ms "exposed all-reduce" of the r2 runs)
            torch._foreach_copy_(grads, views)
        self._inflight = []
        self._pending = [len(b) for b in self.buckets]

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
