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


def _rank_weight(local_clips, global_clips, group=None) -> float:
    if local_clips is None or global_clips is None:
        return 1.0 / dist.get_world_size(group)
    if global_clips <= 0 or local_clips < 0 or local_clips > global_clips:
        raise ValueError(f"bad clip counts: local {local_clips}, global {global_clips}")
    return float(local_clips) / float(global_clips)


class GradientAllReducer:
    """The single collective of a data-parallel training step: every parameter gradient is packed into ONE flat
    fp32 buffer (29.4 M elements for VGG-16 full), all-reduced once (NCCL over NVLink/NVSwitch on
    the GPU box, gloo in the CPU tests), averaged over the world and unpacked in place.  Replaces what
    nn.DataParallel (train_net_dynamic.py:96) does implicitly by gathering replicas' gradients on GPU 0.

        reducer = GradientAllReducer(model.parameters())
        loss.backward(); reducer(local_clips, global_clips); optimizer.step()

    Each rank's loss is the mean over ITS clips, so its gradient enters the sum with weight local_clips / global_clips
    (shard_range hands out unequal -- possibly empty -- shards when the batch does not divide by the world size): the
    result is the global-batch mean gradient, what the reference's single loss over the gathered batch produces.
    Without the two counts every rank weighs 1 / world (equal shards).

    This is the simple post-backward form (one eager copy per parameter each way).  BucketedGradientReducer below is
    the one the training launcher and bench.py use: packed by one kernel launch per bucket and overlapped with the
    backbone's backward.
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def __call__(self, local_clips=None, global_clips=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        weight = _rank_weight(local_clips, global_clips, self.group)
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
        flat.mul_(weight)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)          # the one collective of the step
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        return self.numel


class BucketedGradientReducer:
    """The data-parallel training step's gradient exchange, overlapped with the backward pass.

    The drop-in models run their whole backward inside one autograd node (infer_model._DinTrainFn), which hands its
    gradients to `model.grad_sink` in two groups, in the order the backward produces them:

      "head"      everything after the backbone (fc_emb_1's 13.1 M weights are half of all gradient bytes) -- complete
                  BEFORE the backbone's backward starts;
      "backbone"  the convolution weights, complete at the end.

    Each group is scaled by the rank's weight (local_clips / global_clips) and packed into its region of ONE flat fp32
    buffer by a single kernel launch (din_pack_flat_f32), and its all-reduce is issued immediately with async_op=True:
    NCCL runs it on its own stream, so the head bucket's exchange over NVLink overlaps the backbone's dgrad / wgrad
    kernels, and only the backbone bucket's exchange is exposed at the end of the step.  `finish()` waits for both and
    makes every parameter's `.grad` a VIEW of the flat buffer -- there is no unpack copy.  (Replaces nn.DataParallel's
    per-step replicate + gather + reduce on GPU 0, train_net_dynamic.py:96,220-224; first version: GradientAllReducer.)

        reducer = BucketedGradientReducer(model)          # installs model.grad_sink
        for batch in loader:
            reducer.set_batch(local_clips, global_clips)  # optional: unequal shards
            optimizer.zero_grad(); loss = ...; loss.backward(); optimizer.step()
    """

    STAGES = ("head", "backbone")

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.by_stage = {"head": [(n, p) for n, p in self.named if not n.startswith("backbone.")],
                         "backbone": [(n, p) for n, p in self.named if n.startswith("backbone.")]}
        self.offsets, self.regions, off = {}, {}, 0
        for stage in self.STAGES:
            start = off
            for n, p in self.by_stage[stage]:
                self.offsets[n] = off
                off += (p.numel() + 3) // 4 * 4                # 16-byte aligned slots: float4 packing, aligned views
            self.regions[stage] = (start, off)
        self.numel = off
        self._flat, self._works, self._weight = None, [], None
        self.stats = {"steps": 0, "bytes": 0}
        model.grad_sink = self

    # -- per step -------------------------------------------------------------------------------------------------
    def set_batch(self, local_clips, global_clips):
        self._weight = _rank_weight(local_clips, global_clips, self.group)

    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _buffer(self, device):
        if self._flat is None or self._flat.device != device:
            self._flat = torch.zeros((self.numel,), dtype=torch.float32, device=device)
        return self._flat

    def __call__(self, stage, grads):
        """Called from the backward with the gradient dict once every tensor of `stage` is final."""
        entries = self.by_stage[stage]
        if not entries or not self.active():
            return
        dev = entries[0][1].device
        flat = self._buffer(dev)
        a, b = self.regions[stage]
        weight = self._weight if self._weight is not None else 1.0 / dist.get_world_size(self.group)
        have = [(n, grads[n]) for n, _ in entries if grads.get(n) is not None]
        if len(have) != len(entries):
            flat[a:b].zero_()                                  # a parameter without a gradient contributes zeros
        if have:
            tensors = [g.contiguous().float() for _, g in have]
            offs = [self.offsets[n] for n, _ in have]
            if dev.type == "cuda":
                from . import ops
                ops.pack_flat(tensors, flat, offs, weight)     # one launch for the whole bucket
            else:                                              # host-logic tests (gloo): same layout through torch
                for t, o in zip(tensors, offs):
                    flat[o:o + t.numel()].copy_(t.reshape(-1)).mul_(weight)
        self._works.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.stats["bytes"] += 4 * (b - a)

    def finish(self):
        """Wait for the outstanding all-reduces; every parameter's .grad becomes (or accumulates) its flat view."""
        for w in self._works:
            w.wait()
        self._works = []
        self.stats["steps"] += 1
        flat = self._flat
        for n, p in self.named:
            view = flat[self.offsets[n]:self.offsets[n] + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                p.grad.add_(view)
        self._weight = None
