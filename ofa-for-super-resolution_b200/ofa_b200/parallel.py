"""Multi-GPU plumbing of the SR path: one process per GPU, torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).

Inference: frames — or spatial tiles with a halo that covers the network's receptive
field — are independent units, so they are SHARDED across ranks with NO data-path collective.  In
eval mode BatchNorm is a per-channel affine, so a tile computed with `halo` extra LR pixels on every
interior side and cropped afterwards equals the same region of the whole-frame result exactly.

Training (progressive shrinking): batch-sharded; every rank seeds Python `random` identically
(progressive_shrinking.py:164) so all ranks sample the same sub-network; gradients accumulate over
`dynamic_batch_size` sub-networks and are then summed across ranks ONCE per step through a flat,
zero-filled buffer (inactive blocks have no .grad — a fixed-size buffer keeps the collective
shape-stable), divided by the world size (Horovod semantics, distributed_run_manager.py:72-75).
BatchNorm statistics stay per-rank, as under nn.DataParallel.
"""
import torch

# receptive-field radius of the max S4 sub-network in LR pixels: stem 2 + 14 blocks x 3 (7x7) + tail
# 2 + 2 + shuffle convs 2 + 1 + output conv 0.5  ->  51.5; 64 keeps tiles 16-aligned
S4_HALO_LR = 64


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of `n_items` units for `rank` (frames, tiles or samples)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def tile_grid(h, w, tiles_y, tiles_x, halo):
    """Split an h x w LR frame into tiles_y x tiles_x tiles.  Returns a list of dicts with the input
    window (with halo, clipped to the frame) and the crop that removes the halo again, both in LR
    pixels: {'in': (y0, y1, x0, x1), 'core': (y0, y1, x0, x1), 'crop': (cy0, cy1, cx0, cx1)}."""
    tiles = []
    for ty in range(tiles_y):
        cy0, cy1 = shard_range(h, ty, tiles_y)
        for tx in range(tiles_x):
            cx0, cx1 = shard_range(w, tx, tiles_x)
            iy0, iy1 = max(0, cy0 - halo), min(h, cy1 + halo)
            ix0, ix1 = max(0, cx0 - halo), min(w, cx1 + halo)
            tiles.append({'in': (iy0, iy1, ix0, ix1), 'core': (cy0, cy1, cx0, cx1),
                          'crop': (cy0 - iy0, cy1 - iy0, cx0 - ix0, cx1 - ix0)})
    return tiles


def tiled_forward(net, x, tiles_y, tiles_x, halo=S4_HALO_LR, scale=4, rank=0, world=1, out=None):
    """Run `net` on this rank's share of the tile grid of frame batch `x` [N,3,h,w].  Returns
    (out, my_tiles): `out` is an [N,3,scale*h,scale*w] tensor in which only this rank's core regions are
    written (no collective — the caller owns assembly, e.g. each rank writes its region of a shared
    file / display buffer)."""
    n, _, h, w = x.shape
    tiles = tile_grid(h, w, tiles_y, tiles_x, halo)
    lo, hi = shard_range(len(tiles), rank, world)
    if out is None:
        out = torch.zeros((n, 3, scale * h, scale * w), dtype=torch.float32, device=x.device)
    for t in tiles[lo:hi]:
        iy0, iy1, ix0, ix1 = t['in']
        y = net(x[:, :, iy0:iy1, ix0:ix1])
        cy0, cy1, cx0, cx1 = (scale * v for v in t['crop'])
        oy0, oy1, ox0, ox1 = (scale * v for v in t['core'])
        out[:, :, oy0:oy1, ox0:ox1] = y[:, :, cy0:cy1, cx0:cx1]
    return out, tiles[lo:hi]


class FlatGradAllReduce:
    """Gradient exchange of data-parallel progressive-shrinking training.

    All parameters own a slot in one flat fp32 buffer (fixed layout, so every rank issues the same
    collectives whatever sub-network was sampled; inactive blocks simply contribute zeros).  The buffer is
    split into a TAIL segment (parameters of the layers that run last in forward, whose gradients are
    complete first in backward) and a HEAD segment.

    * `reduce()` alone: packs the available .grad tensors, all-reduces the segments (tail first) in
      `n_buckets` buckets each, averages, and scatters the result back into the existing .grad tensors
      (inactive parameters have no .grad on any rank and keep none: the optimizer skips them).
    * overlapped with backward: `watch(boundary_tensor)` before the step's LAST `loss.backward()` registers a
      gradient hook on the activation that separates tail from head.  When autograd reaches it every tail
      gradient is final, so the tail segment is packed and its all-reduce is launched asynchronously while
      the head's backward is still running; `reduce()` then only has the head segment left to send.
    """

    def __init__(self, params, n_buckets=2, process_group=None, tail_params=None):
        params = [p for p in params if p.requires_grad]
        tail_ids = {id(p) for p in (tail_params or [])}
        self.tail = [p for p in params if id(p) in tail_ids]
        self.head = [p for p in params if id(p) not in tail_ids]
        self.params = self.tail + self.head            # flat layout: tail segment first
        self.group = process_group
        self.offsets = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += p.numel()
        self.total = total
        self.tail_numel = sum(p.numel() for p in self.tail)
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views = [self.flat[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets)]
        self.n_buckets = n_buckets
        self._tail_handles = None

    def _buckets(self, lo, hi):
        n = hi - lo
        cuts = [lo + round(n * i / self.n_buckets) for i in range(self.n_buckets + 1)]
        return [(cuts[i], cuts[i + 1]) for i in range(self.n_buckets) if cuts[i + 1] > cuts[i]]

    def _pack(self, first, last):
        """Gradients of params[first:last] -> their slots of the flat buffer: one zero fill of the segment (inactive
        blocks contribute zeros) and ONE multi-tensor copy of the gradients that exist (a per-parameter copy loop
        costs ~10 us of host time per parameter: more than the whole all-reduce)."""
        if last <= first:
            return
        lo = self.offsets[first]
        hi = self.offsets[last - 1] + self.params[last - 1].numel()
        self.flat[lo:hi].zero_()
        dst = [v for p, v in zip(self.params[first:last], self.views[first:last]) if p.grad is not None]
        src = [p.grad for p in self.params[first:last] if p.grad is not None]
        if dst:
            torch._foreach_copy_(dst, src)

    def _launch(self, lo, hi):
        import torch.distributed as dist
        return [dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                for a, b in self._buckets(lo, hi)]

    def watch(self, boundary):
        """Arm the overlap for the next backward: `boundary` is the activation tensor between the head layers
        and the `tail_params` layers (it must require grad)."""
        assert self.tail, 'FlatGradAllReduce(..., tail_params=...) is needed for the overlapped mode'

        def _hook(grad):
            if self._tail_handles is None:
                self._pack(0, len(self.tail))
                self._tail_handles = self._launch(0, self.tail_numel)
            return grad
        boundary.register_hook(_hook)

    def reduce(self):
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        handles = self._tail_handles
        if handles is None:                            # not overlapped: send the tail now
            self._pack(0, len(self.tail))
            handles = self._launch(0, self.tail_numel) if self.tail_numel else []
        self._tail_handles = None
        self._pack(len(self.tail), len(self.params))
        handles += self._launch(self.tail_numel, self.total)
        for h in handles:
            h.wait()
        self.flat.div_(world)
        # every rank sampled the same sub-network (identical `random` seeds), so a parameter without a gradient has
        # none on any rank: it stays None and the optimizer skips it, exactly as in the single-GPU reference
        dst = [p.grad for p in self.params if p.grad is not None]
        src = [v for p, v in zip(self.params, self.views) if p.grad is not None]
        if dst:
            torch._foreach_copy_(dst, src)


def s4_tail_parameters(net):
    """Tail / head split of OFAMobileNetS4 for the overlapped all-reduce: everything after the trunk's long skip
    (dec_final_conv_blocks, the PixelShuffle stages at the end of `blocks`, the output conv) is the tail; the
    boundary activation is the input of dec_final_conv_blocks[0].  Returns (tail_params, boundary_module)."""
    tail = list(net.dec_final_conv_blocks.parameters()) + list(net.dec_final_output_conv_block.parameters())
    n_shuffle = len(net.block_group_info) - 4
    for group in net.block_group_info[len(net.block_group_info) - n_shuffle:]:
        for idx in group:
            tail += list(net.blocks[idx].parameters())
    return tail, net.dec_final_conv_blocks[0]


def broadcast_parameters(module, src=0, process_group=None):
    """One-time broadcast of parameters and buffers from rank `src` (distributed_run_manager.py:180-184)."""
    import torch.distributed as dist
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
    from . import functional as OF
    OF.invalidate_packed_weights()          # .data writes do not move Tensor._version
