"""parallel.py — multi-GPU plumbing for the DIN path: one process per GPU, clips sharded, no collective on the
forward data path; training adds ONE all-reduce over a flat gradient buffer per step.

Replaces the reference's nn.DataParallel (train_net_dynamic.py:95-96), which re-broadcasts all parameters
and scatters / gathers activations through GPU 0 on every step.  Here every rank holds a replica of the
weights, takes a contiguous shard of the clips (every clip is independent in the forward pass, SURVEY.md
§8e), and only the [B, num_activities] logits are gathered (eval) — via torch.distributed (NCCL on the GPU
box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Balanced contiguous split of range(n_items): the first (n_items % world) ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch, rank: int, world: int):
    """Slice every tensor of a (images, boxes[, bboxes_num]) tuple along the clip dimension."""
    n = batch[0].shape[0]
    a, b = shard_range(n, rank, world)
    return tuple(t[a:b] for t in batch)


def all_gather_logits(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Every rank contributes the logits of its shard (possibly empty); returns [n_total, A] in clip order
    on every rank.  Shards are padded to the largest shard so that a plain all_gather suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_len = max(b - a for a, b in sizes)
    a_dim = local.shape[1]
    a0, b0 = sizes[rank]
    if local.shape[0] != b0 - a0:
        raise ValueError(f"rank {rank}: expected {b0 - a0} rows, got {local.shape[0]}")
    padded = local.new_zeros((max_len, a_dim))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-measured time."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def sharded_forward(model, batch, n_total=None, group=None):
    """Evaluate `model` on this rank's shard of `batch` (full batch given on every rank) and return the
    gathered logits for all clips.  The forward itself involves no communication."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_total = batch[0].shape[0] if n_total is None else n_total
    local = shard_batch(batch, rank, world)
    if local[0].shape[0] > 0:
        with torch.no_grad():
            out = model(local)["activities"]
    else:
        out = batch[0].new_zeros((0, model.cfg.num_activities), dtype=torch.float32)
    return all_gather_logits(out, n_total, group)


class GradientAllReducer:
    """The single collective of a data-parallel training step: every parameter gradient is packed into ONE flat
    fp32 buffer (29.4 M elements for VGG-16 full), all-reduced once (NCCL over NVLink/NVSwitch on
    the GPU box, gloo in the CPU tests), averaged over the world and unpacked in place.  Replaces what
    nn.DataParallel (train_net_dynamic.py:96) does implicitly by gathering replicas' gradients on GPU 0.

        reducer = GradientAllReducer(model.parameters())
        loss.backward(); reducer(); optimizer.step()
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        dev = self.params[0].device
        if self._flat is None or self._flat.device != dev:
            self._flat = torch.zeros((self.numel,), dtype=torch.float32, device=dev)
        flat, off = self._flat, 0
        for p in self.params:                       # parameters without a gradient this step contribute zeros
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)          # the one collective of the step
        flat.div_(dist.get_world_size(self.group))
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        return self.numel
