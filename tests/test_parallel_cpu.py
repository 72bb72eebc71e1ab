"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: unit sharding, the halo tile grid and the
flat gradient all-reduce of data-parallel progressive shrinking.  The kernels are not involved."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ofa_b200 import parallel as P


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [P.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_tile_grid_covers_frame_once_and_halo_is_clipped():
    h, w, halo = 540, 960, 64
    tiles = P.tile_grid(h, w, 2, 4, halo)
    cover = torch.zeros(h, w, dtype=torch.int32)
    for t in tiles:
        iy0, iy1, ix0, ix1 = t['in']
        cy0, cy1, cx0, cx1 = t['core']
        assert 0 <= iy0 <= cy0 < cy1 <= iy1 <= h and 0 <= ix0 <= cx0 < cx1 <= ix1 <= w
        assert (cy0 - iy0 in (0, halo)) and (cx0 - ix0 in (0, halo))      # halo only on interior sides
        assert t['crop'] == (cy0 - iy0, cy1 - iy0, cx0 - ix0, cx1 - ix0)
        cover[cy0:cy1, cx0:cx1] += 1
    assert int(cover.min()) == 1 and int(cover.max()) == 1


def test_tiled_forward_equals_whole_frame_for_a_local_operator():
    """A stand-in 'network' with a known receptive field (5 stacked 3x3 box filters + 4x nearest
    upsampling): tiling with halo >= radius must reproduce the whole-frame result exactly."""
    def net(x):
        for _ in range(5):
            x = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (1, 1, 1, 1)), 3, 1) * 9
        return torch.nn.functional.interpolate(x, scale_factor=4, mode='nearest')
    x = torch.rand(2, 3, 37, 53)
    whole = net(x)
    out = None
    for rank in range(3):
        out, mine = P.tiled_forward(net, x, 2, 3, halo=5, scale=4, rank=rank, world=3, out=out)
        assert len(mine) == 2
    assert torch.equal(out, whole)


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # identical init on every rank
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
        if rank == 1:                              # diverge, then broadcast from rank 0 must repair it
            for p in model.parameters():
                p.data.add_(1.0)
        P.broadcast_parameters(model, src=0)
        # BN re-calibration meter (elastic_nn/utils.py DistributedTensor): batch-weighted local sums, ONE all-reduce
        from ofa_b200.elastic_nn.utils import _Avg
        meter = _Avg()
        meter.update(torch.full((4,), float(rank + 1)), 3)
        meter.update(torch.full((4,), float(10 * (rank + 1))), 2)
        expect = sum((3.0 * (r + 1) + 2.0 * 10 * (r + 1)) / 5.0 for r in range(world)) / world
        assert torch.allclose(meter.avg(True), torch.full((4,), expect))
        sync = P.FlatGradAllReduce(model.parameters(), n_buckets=2)
        # every rank seeds `random` identically -> same "sub-network" choice (progressive_shrinking.py:164)
        random.seed(int('%d%.3d%.3d' % (7, 0, 0)))
        skip_last = random.choice([True, False])
        torch.manual_seed(100 + rank)              # different data shard per rank
        x = torch.randn(8, 6)
        h = model[1](model[0](x))
        y = h.sum() if skip_last else model[2](h).sum()   # the skipped layer has NO .grad on any rank
        y.backward()
        local = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
        sync.reduce()
        gathered = [None] * world
        dist.all_gather_object(gathered, [None if g is None else g.tolist() for g in local])
        ok = True
        for i, p in enumerate(model.parameters()):
            parts = [torch.tensor(g[i]) if g[i] is not None else torch.zeros_like(p) for g in gathered]
            expect = sum(parts) / world
            if all(g[i] is None for g in gathered):
                ok = ok and p.grad is None                # inactive everywhere: stays None, the optimizer skips it
            else:
                ok = ok and p.grad is not None and torch.allclose(p.grad, expect, atol=1e-6)
        w0 = [p.detach().clone() for p in model.parameters()]
        allw = [None] * world
        dist.all_gather_object(allw, [w.tolist() for w in w0])
        ok = ok and allw[0] == allw[1]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _worker_overlap(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        head = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 5))
        unused = torch.nn.Linear(5, 5)                 # an inactive block: no .grad on any rank
        tail = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
        params = list(head.parameters()) + list(unused.parameters()) + list(tail.parameters())
        sync = P.FlatGradAllReduce(params, n_buckets=2, tail_params=list(tail.parameters()))
        launched_in_backward = []
        orig_launch = sync._launch

        def spy(lo, hi):
            launched_in_backward.append((lo, hi, all(p.grad is None for p in head.parameters())))
            return orig_launch(lo, hi)
        sync._launch = spy
        torch.manual_seed(100 + rank)
        x = torch.randn(8, 6)
        h = head(x)
        sync.watch(h)                                  # boundary activation between head and tail
        tail(h).pow(2).sum().backward()
        local = [None if p.grad is None else p.grad.clone() for p in params]
        sync.reduce()
        gathered = [None] * world
        dist.all_gather_object(gathered, [None if g is None else g.tolist() for g in local])
        ok = True
        for i, p in enumerate(params):
            parts = [torch.tensor(g[i]) if g[i] is not None else torch.zeros_like(p) for g in gathered]
            if all(g[i] is None for g in gathered):
                ok = ok and p.grad is None
            else:
                ok = ok and p.grad is not None and torch.allclose(p.grad, sum(parts) / world, atol=1e-6)
        # the tail segment went out while the head had no gradients yet (i.e. during backward), the rest after
        ok = ok and len(launched_in_backward) == 2 and launched_in_backward[0] == (0, sync.tail_numel, True)
        ok = ok and launched_in_backward[1][:2] == (sync.tail_numel, sync.total) and launched_in_backward[1][2] is False
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_overlaps_backward_world2_gloo():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_overlap, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_flat_grad_allreduce_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
