"""A stage-2 training loop on the CUDA path with synthetic data -- the loop of train_net_dynamic.py:157-236 with the
reference's model class swapped for the drop-in, F.cross_entropy + the per-step `.item()` syncs replaced by
din_b200.metrics (one launch, meters stay on the device), and nn.DataParallel replaced by one process per GPU with a
single flat gradient all-reduce per step.

  python tests/tools/train_stage2_synthetic.py --steps 20                      # one GPU
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/train_stage2_synthetic.py
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import din_oracle as O  # noqa: E402  (synthetic weights / inputs only: the oracle's generators)
import infer_model as IM  # noqa: E402
from config import Config  # noqa: E402
from din_b200 import metrics  # noqa: E402
from din_b200.parallel import GradientAllReducer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=2, help="clips per GPU (scripts/train_volleyball_stage2_dynamic.py:42)")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--freeze-backbone", action="store_true", help="config.py:39 default (train_backbone = False)")
    ap.add_argument("--lr", type=float, default=1e-4)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    hw = (args.height, args.width)
    pc = O.PathConfig(backbone="vgg16", image_size=hw, out_size=O.backbone_out_size("vgg16", *hw),
                      num_frames=args.frames, num_boxes=12)
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "lite_dim",
              "ST_kernel_size", "scale_factor", "beta_factor", "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    cfg.train_backbone = not args.freeze_backbone
    model = IM.Dynamic_volleyball(cfg)
    model.load_state_dict(O.make_state_dict(pc, seed=0))
    model = model.to(dev).train()
    params = [q for q in model.parameters() if q.requires_grad]
    optimizer = torch.optim.Adam(params, lr=args.lr, weight_decay=cfg.weight_decay)
    reducer = GradientAllReducer(params)
    meters = metrics.DeviceMeters(cfg.num_activities, dev)
    images, boxes = (t.to(dev) for t in O.make_inputs(pc, args.batch, seed=rank))
    labels = (torch.arange(args.batch, device=dev) + rank) % cfg.num_activities
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(args.steps):
        scores = model((images, boxes))["activities"]
        loss = metrics.cross_entropy(scores, labels, meters=meters)       # loss + accuracy + confusion: one launch
        optimizer.zero_grad()
        loss.backward()
        reducer()                                                         # the step's only collective
        optimizer.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = meters.value()                                                 # the only device->host read of the run
    if rank == 0:
        print(f"{args.steps} steps x {args.batch * world} clips: {args.steps * args.batch * world / dt:.1f} clips/s; "
              f"mean loss {info['loss']:.4f}, accuracy {info['activities_acc']:.1f} %, last loss {loss.item():.4f}")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
