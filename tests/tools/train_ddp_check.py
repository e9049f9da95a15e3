"""Data-parallel training step over NCCL (run under torchrun, one rank per GPU):
every rank runs forward + backward (VGG-16 backbone trained) on its shard of the clips; the gradients are packed into
one flat buffer in two buckets and all-reduced while the backward still runs (din_b200.parallel.BucketedGradientReducer,
shards of UNEQUAL size weighted by local / global clips), and the result is compared with the gradient of the same
global batch computed on a single GPU (rank 0).  Dropout is off so that the two are comparable.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/train_ddp_check.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import din_oracle as O  # noqa: E402  (synthetic weights / inputs only)
import infer_model as IM  # noqa: E402
from config import Config  # noqa: E402
from din_b200 import metrics  # noqa: E402
from din_b200.parallel import BucketedGradientReducer, shard_batch, shard_range  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)

hw = (360, 640)
pc = O.PathConfig(backbone="vgg16", image_size=hw, out_size=O.backbone_out_size("vgg16", *hw), num_frames=4, num_boxes=12)
sd = O.make_state_dict(pc, seed=0)
B = 2 * world + 1                                # does not divide by the world size: unequal shards
batch = O.make_inputs(pc, B, seed=0)
labels = torch.arange(B) % pc.num_activities


def build():
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "lite_dim",
              "ST_kernel_size", "scale_factor", "beta_factor", "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    cfg.train_backbone, cfg.train_dropout_prob = True, 0.0
    m = IM.Dynamic_volleyball(cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).train()


def step(model, clips, lab):
    for q in model.parameters():
        q.grad = None
    out = model(tuple(t.to(dev) for t in clips))["activities"]
    loss = metrics.cross_entropy(out, lab.to(dev))
    loss.backward()
    return loss


model = build()
a, b = shard_range(B, rank, world)
local = shard_batch(batch, rank, world)
reducer = BucketedGradientReducer(model)        # installs model.grad_sink: the exchange happens inside backward()
reducer.set_batch(b - a, B)
step(model, local, labels[a:b])                 # warm-up (weight packing, allocator, NCCL channels)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reducer.set_batch(b - a, B)
loss = step(model, local, labels[a:b])
n = reducer.numel
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ref = build()
    step(ref, batch, labels)
    worst = 0.0
    for (k, q), (_, r) in zip(model.named_parameters(), ref.named_parameters()):
        err = float((q.grad - r.grad).double().norm() / max(float(r.grad.double().norm()), 1e-30))
        worst = max(worst, err)
    print(f"world {world}: {B} clips, {n} gradient elements in 2 overlapped all-reduces, step {float(ms):.2f} ms (max over "
          f"ranks); worst relative L2 difference to the single-GPU gradient of the same global batch: {worst:.2e}")
    assert worst < 2e-3, worst      # fp32 atomics / summation order only
dist.barrier()
dist.destroy_process_group()
