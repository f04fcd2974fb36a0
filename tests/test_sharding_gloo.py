"""Host-side sharding / gather logic of the N > 1 path on CPU: world_size 2 (and 3, uneven) over gloo.
The model is a stand-in with the module's call signature; the kernels themselves are covered by -m gpu."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mhada_style_transfer_b200.sharding import gather_batches, run_sharded, shard_bounds, shard_range


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 9, 64):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(9, 0, 2) == (0, 5) and shard_range(9, 1, 2) == (5, 9)
    with pytest.raises(ValueError):
        shard_bounds(4, 0)


def _fake_model(fc, fs):
    # per-image function of the inputs (like MHAda + decoder: no coupling across the batch)
    fcs = fc[0] * 2.0 + fs[0].mean(dim=(1, 2, 3), keepdim=True)
    cs = torch.stack([fcs[:, :3].sum(dim=(1, 2, 3))] * 4, dim=1).reshape(-1, 1, 2, 2) + fc[1][:, :1, :2, :2]
    return fcs, cs


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    fc = [torch.randn(n_images, 8, 4, 4, generator=g) for _ in range(3)]
    fs = [torch.randn(n_images, 8, 3, 3, generator=g) for _ in range(3)]
    got = run_sharded(_fake_model, fc, fs, dst=0)
    if rank == 0:
        want = _fake_model(fc, fs)[1]
        q.put((tuple(got.shape), bool(torch.equal(got, want))))
    else:
        assert got is None
    # an uneven explicit gather too
    s, e = shard_range(n_images, rank, world)
    full = torch.arange(n_images * 3, dtype=torch.float32).reshape(n_images, 3)
    out = gather_batches(full[s:e], n_images, dst=0)
    if rank == 0:
        q.put(bool(torch.equal(out, full)))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_images", [(2, 8), (2, 5), (3, 2)])
def test_sharded_run_matches_unsharded(world, n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    shape, same = q.get(timeout=10)
    assert shape == (n_images, 1, 2, 2) and same
    assert q.get(timeout=10) is True


def _overlap_worker(rank, world, port, q):
    from mhada_style_transfer_b200.sharding import OverlappedGradientAllReduce, allreduce_gradients
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4))
    unused = torch.nn.Linear(3, 3)                       # a module whose parameters get no gradient
    ref = [p.detach().clone() for p in net.parameters()]
    x = torch.randn(8, 16, generator=torch.Generator().manual_seed(1))
    s, e = shard_range(8, rank, world)
    sync = OverlappedGradientAllReduce([net, unused], bucket_bytes=256)        # several buckets
    assert len(sync.buckets) >= 3
    for step in range(2):                                # two steps: the hooks re-arm
        for p in net.parameters():
            p.grad = None
        (net(x[s:e]).pow(2).sum() / 8 * world).backward()   # shard loss scaled so that the average = full-batch gradient
        launched = sync.launched_in_backward
        sync.finish()
    got = [p.grad.clone() for p in net.parameters()]
    # the plain (post-backward) all-reduce gives the same averages
    for p in net.parameters():
        p.grad = None
    (net(x[s:e]).pow(2).sum() / 8 * world).backward()
    sync.remove()
    allreduce_gradients(net)
    same = all(torch.allclose(a, p.grad, rtol=1e-6, atol=1e-7) for a, p in zip(got, net.parameters()))
    # and both equal the full-batch gradient
    full = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                               torch.nn.Linear(32, 4))
    with torch.no_grad():
        for p, r in zip(full.parameters(), ref):
            p.copy_(r)
    (full(x).pow(2).sum() / 8).backward()
    close = all(torch.allclose(a, p.grad, rtol=1e-5, atol=1e-6) for a, p in zip(got, full.parameters()))
    if rank == 0:
        q.put((same, close, launched >= 2, unused.weight.grad is None))
    dist.destroy_process_group()


def test_overlapped_gradient_allreduce_matches_full_batch():
    """BASELINE configs[4]: gradient all-reduce launched bucket by bucket from autograd hooks DURING backward."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == (True, True, True, True)


def _disabled_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    from mhada_style_transfer_b200.sharding import OverlappedGradientAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Linear(8, 4)
    sync = OverlappedGradientAllReduce(net, bucket_bytes=64)
    x = torch.full((2, 8), float(rank + 1))
    sync.enabled = False                       # hooks and finish() do nothing: gradients stay local
    net(x).sum().backward()
    sync.finish()
    local = net.weight.grad.clone()
    net.zero_grad(set_to_none=True)
    sync.enabled = True
    net(x).sum().backward()
    sync.finish()
    q.put((rank, float(local[0, 0]), float(net.weight.grad[0, 0])))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_can_be_disabled_for_timing():
    """bench.py times the captured training step without its collectives: `enabled = False` leaves the gradients local,
    re-enabling restores the averaging."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_disabled_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got[0][1] == 2.0 and got[1][1] == 4.0          # local gradients: 2 rows x (rank + 1)
    assert got[0][2] == got[1][2] == 3.0                   # averaged over the two ranks
