"""Runs the reference's UNMODIFIED experiment script `scripts/train_volleyball_stage2_dynamic.py` -- and through it the
unmodified `train_net_dynamic.train_net` (reference train_net_dynamic.py:27-157) -- against this package's drop-in
modules, exactly as INTEGRATION.md §1 documents: the package directory FIRST on sys.path, the reference root after it.

What the harness adds around the two unmodified files (and nothing inside them):
  * `return_dataset` (dataset.py:7, needs the Volleyball videos) is replaced by a synthetic dataset with the loader's
    item contract (volleyball.py:270-276: images [T,3,H,W] float 0..255, boxes [T,N,4] in feature-map units,
    actions [T,N] long, activities [T] long);
  * `train_net` is called through a wrapper that only shortens the run (max_epoch, device_list);
  * the stage-1 checkpoint the script loads (`cfg.stage1_model_path`) is written beforehand by this package's
    `Basenet_volleyball.savemodel` -- the stage-1 -> stage-2 hand-over of the real recipe;
  * an empty `skimage` stub (volleyball.py:2-3 imports it and never uses it) when scikit-image is not installed.

    python tests/tools/dropin_run.py --ref <reference root> --workdir <scratch dir> [--epochs 1] [--devices 0]
Prints one JSON line: losses parsed from the trainer's own log, the checkpoint it wrote, plan-build counts.
"""
import argparse
import glob
import json
import os
import re
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--train-clips", type=int, default=4)
    ap.add_argument("--test-clips", type=int, default=1)
    ap.add_argument("--script", default="train_volleyball_stage2_dynamic.py",
                    help="the reference experiment script under scripts/ (also: train_volleyball_stage2_dynamic_tce.py)")
    ap.add_argument("--import-only", action="store_true", help="resolve every name the trainer needs, then stop (CPU)")
    args = ap.parse_args()
    os.environ.setdefault("DIN_OFFLINE", "1")
    ref = os.path.abspath(args.ref)
    sys.path[:0] = [PKG, ref]                                   # the documented order: package first, reference second
    try:
        import skimage  # noqa: F401
    except ImportError:
        sys.path.append(os.path.join(ROOT, "oracle", "shims"))  # empty skimage stub; LAST, so it shadows nothing
    os.makedirs(os.path.join(args.workdir, "result"), exist_ok=True)
    os.chdir(args.workdir)

    import torch
    import train_net_dynamic as T                               # the reference's file, byte for byte
    assert os.path.dirname(os.path.abspath(T.__file__)) == ref, T.__file__
    import infer_model
    import utils
    assert os.path.dirname(os.path.abspath(infer_model.__file__)) == PKG, infer_model.__file__
    assert os.path.dirname(os.path.abspath(utils.__file__)) == PKG, utils.__file__
    # every global name train_net_dynamic.py uses must resolve in its namespace (the star-imports deliver them)
    import builtins
    import dis
    code_names = set()

    def walk(code):
        for ins in dis.get_instructions(code):
            if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME"):
                code_names.add(ins.argval)
        for c in code.co_consts:
            if hasattr(c, "co_code"):
                walk(c)
    with open(T.__file__) as fh:
        walk(compile(fh.read(), T.__file__, "exec"))
    missing = sorted(n for n in code_names if not hasattr(T, n) and not hasattr(builtins, n))
    assert not missing, f"names the reference trainer uses but the drop-in does not provide: {missing}"
    if args.import_only:
        print(json.dumps({"resolved": len(code_names), "missing": missing}))
        return

    class SyntheticVolleyball(T.data.Dataset):
        def __init__(self, n, cfg, seed):
            self.n, self.cfg, self.seed = n, cfg, seed

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            c = self.cfg
            g = torch.Generator().manual_seed(self.seed * 1000 + i)
            t, n = c.num_frames, c.num_boxes
            oh, ow = c.out_size
            images = torch.randint(0, 256, (t, 3) + tuple(c.image_size), generator=g, dtype=torch.uint8).float()
            cx, cy = torch.rand(t, n, generator=g) * ow, torch.rand(t, n, generator=g) * oh
            w, h = 1 + 3 * torch.rand(t, n, generator=g), 2 + 5 * torch.rand(t, n, generator=g)
            boxes = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), dim=-1)
            actions = torch.randint(0, c.num_actions, (t, n), generator=g)
            activities = torch.full((t,), int(torch.randint(0, c.num_activities, (1,), generator=g)))
            return images, boxes, actions, activities

    T.return_dataset = lambda cfg: (SyntheticVolleyball(args.train_clips, cfg, 1),
                                    SyntheticVolleyball(args.test_clips, cfg, 2))
    real_train_net = T.train_net
    seen = {}

    def short_train_net(cfg):
        cfg.max_epoch, cfg.device_list = args.epochs, args.devices
        # the stage-1 checkpoint the script points at, written by the drop-in stage-1 model
        import base_model
        cfg_log, cfg.log_path = getattr(cfg, "log_path", None), None
        stage1 = base_model.Basenet_volleyball(cfg)
        stage1.savemodel(cfg.stage1_model_path)
        del stage1
        if cfg_log is not None:
            cfg.log_path = cfg_log
        seen["cfg"] = cfg
        return real_train_net(cfg)

    T.train_net = short_train_net
    from din_b200 import plan_cache
    runpy.run_path(os.path.join(ref, "scripts", args.script), run_name="__main__")

    cfg = seen["cfg"]
    log = open(cfg.log_path).read()
    losses = [float(x) for x in re.findall(r"Loss: ([-+0-9.eE]+|nan|inf)", log)]
    ckpts = sorted(glob.glob(os.path.join(glob.escape(cfg.result_path), "stage2_epoch*.pth")))   # the name holds [ ]
    assert ckpts, "train_net wrote no checkpoint"
    state = torch.load(ckpts[-1], map_location="cpu")
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state["state_dict"].items()}
    cfg_log, cfg.log_path = cfg.log_path, None
    cls = infer_model.Dynamic_TCE_volleyball if cfg.inference_module_name == "dynamic_tce_volleyball" else \
        infer_model.Dynamic_volleyball
    fresh = cls(cfg)
    fresh.load_state_dict(sd, strict=True)
    cfg.log_path = cfg_log
    fresh = fresh.cuda().eval()
    images, boxes, _, _ = SyntheticVolleyball(1, cfg, 2)[0]
    with torch.no_grad():
        logits = fresh((images[None].cuda(), boxes[None].cuda()))["activities"]
    print(json.dumps({"losses": losses, "checkpoint": os.path.basename(ckpts[-1]), "epochs": state["epoch"],
                      "reloaded_logits_finite": bool(torch.isfinite(logits).all()),
                      "data_parallel": bool(cfg.use_multi_gpu), "visible_gpus": torch.cuda.device_count(),
                      "model": cls.__name__,
                      "optimizer_state_tensors": len(state["optimizer"]["state"])}))


if __name__ == "__main__":
    main()
