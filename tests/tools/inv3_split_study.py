"""Which Inception-v3 layers need split (hi + lo) fp16 weights for the 1e-3 logits bar?  Prints the logits error of
three seeds at 139x203 and one 720p case, and the forward time of 16 frames at 720p, per DIN_INV3_SPLIT setting."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import din_oracle as O  # noqa: E402
import infer_model as IM  # noqa: E402
from config import Config  # noqa: E402

dev = torch.device("cuda:0")


def build(pc, sd):
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "lite_dim",
              "ST_kernel_size", "scale_factor", "beta_factor", "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    m = IM.Dynamic_volleyball(cfg)
    m.load_state_dict(sd)
    return m.to(dev).eval()


def pc_of(hw, T, N):
    return O.PathConfig(backbone="inv3", image_size=hw, out_size=O.backbone_out_size("inv3", *hw), emb_features=1056,
                        num_frames=T, num_boxes=N, lite_dim=None)


cases = [(pc_of((139, 203), 2, 4), 2, s) for s in (0, 1, 2)] + [(pc_of((720, 1280), 2, 12), 1, 0)]
refs = []
for pc, B, seed in cases:
    bb = O.build_backbone("inv3")
    sd = O.make_state_dict(pc, seed=seed, backbone=bb)
    O.load_backbone(bb, sd)
    batch = O.make_inputs(pc, B, seed=seed)
    refs.append((pc, sd, batch, O.volleyball_forward(bb, sd, pc, *batch)))
SETTINGS = ["all", "none", "Conv2d_", "Mixed_5", "Mixed_6", "Conv2d_,Mixed_5"]
for setting in SETTINGS:
    os.environ["DIN_INV3_SPLIT"] = setting
    errs = []
    for pc, sd, batch, ref in refs:
        m = build(pc, sd)
        with torch.no_grad():
            out = m(tuple(t.to(dev) for t in batch))["activities"].cpu()
        errs.append((out - ref).abs().max().item() / ref.abs().max().item())
    pc = pc_of((720, 1280), 8, 12)
    m = build(pc, O.make_state_dict(pc, seed=0))
    batch = tuple(t.to(dev) for t in O.make_inputs(pc, 2, seed=0))
    with torch.no_grad():
        for _ in range(2):
            m(batch)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            m(batch)
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"split={setting:18s} rel logits error {['%.2e' % e for e in errs]}   16 frames 720p: {ms:.2f} ms", flush=True)
