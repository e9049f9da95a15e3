#!/usr/bin/env python
"""bench.py — clips/s of the DIN stage-2 forward on synthetic Volleyball-shaped input.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one forward of the hot path (prep -> VGG-16 -> RoIAlign -> embedding -> Dynamic Relation /
Dynamic Walk -> read-out) over one batch of B clips per GPU.  Prints ONE JSON line (rank 0):
  value      clips/s, whole job, inputs already resident in HBM (CUDA events, barrier + sync both sides,
             max over ranks)
  e2e        the same metric through the public API `model((images, boxes))` with HOST (pinned) inputs:
             every step's host->device copy and the device->host read of its logits are inside the timed
             region (copies are double-buffered on a side stream so they overlap the previous step)
  roofline   the dominant kernel (the tcgen05 implicit-GEMM convolution): algorithmic FLOPs of its launches
             / their CUDA-event durations, against the measured bf16 tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own infer_model classes (staged copy oracle/_ref; else the CPU port under oracle/),
             torch-CPU fp32, all host cores, on a bounded sample of the same workload
  train_step (supplementary) one stage-2 training step on 2 clips per GPU, gradient all-reduce inside at N > 1:
             ms_per_step = the MEDIAN of the timed steps (max over ranks), ms_per_step_mean and step_ms beside it
`--impl reference` times that CPU implementation alone, on the same config / metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("DIN_OFFLINE", "1")     # synthetic weights: never try to download ImageNet checkpoints
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "clips_per_sec"
UNIT = "clips/s"

# workload name -> (PathConfig kwargs, clips per GPU, algorithmic GFLOP per clip [SURVEY.md §8d, BASELINE.md §2])
WORKLOADS = {
    # the reference's own stage-2 recipe (scripts/train_volleyball_stage2_dynamic.py): the configuration the
    # metric "clips/sec (Volleyball T=10 N=12 720p)" is quoted on
    "volleyball_vgg16_lite128_T10_N12_720p": (
        dict(backbone="vgg16", image_size=(720, 1280), out_size=(22, 40), emb_features=512, num_frames=10,
             num_boxes=12, lite_dim=128, ST_kernel_size=[(3, 3)], sampling_ratio=(1,)), 8, 5640.7),
    "volleyball_res18_lite128_T10_N12_720p": (
        dict(backbone="res18", image_size=(720, 1280), out_size=(23, 40), emb_features=512, num_frames=10,
             num_boxes=12, lite_dim=128, ST_kernel_size=[(3, 3)], sampling_ratio=(1,)), 32, 672.8),
    # BASELINE.json configs[1]: Inception-v3, C = 1024 (no lite branch), B = 8 (patched oracle, SURVEY.md §8c bug I)
    "volleyball_inv3_full_T10_N12_720p": (
        dict(backbone="inv3", image_size=(720, 1280), out_size=(87, 157), emb_features=1056, num_frames=10,
             num_boxes=12, lite_dim=None, ST_kernel_size=[(3, 3)], sampling_ratio=(1,)), 8, 1068.0),
    # BASELINE.json configs[3]: ST-factorized DIN, ST_kernel_size [(1,3),(3,1)], hierarchical (C = 1024: hier_LN is
    # hard-coded to [10,12,1024], dynamic_infer_module.py:475); patched oracle, SURVEY.md §8c bug H
    "volleyball_vgg16_hier_st_T10_N12_720p": (
        dict(backbone="vgg16", image_size=(720, 1280), out_size=(22, 40), emb_features=512, num_frames=10,
             num_boxes=12, lite_dim=None, ST_kernel_size=[(1, 3), (3, 1)], sampling_ratio=(1,),
             hierarchical_inference=True), 8, 5641.2),
    # BASELINE.json configs[4]: Collective stage-2 DIN (scripts/train_collective_stage2_dynamic.py): ResNet-18 at
    # 480x720, up to 13 actors per clip (a different count per clip), 4 activities; B = 16 global on 4 GPUs = 4 clips
    # per GPU there (use --global-clips 16 for that strong-scaling shape); patched oracle, bug C
    "collective_res18_T10_N13_480p": (
        dict(dataset="collective", backbone="res18", image_size=(480, 720), out_size=(15, 23), emb_features=512,
             num_frames=10, num_boxes=13, lite_dim=None, ST_kernel_size=(3, 3), sampling_ratio=(1,),
             num_activities=4), 16, 254.8),
    # the sibling model of SURVEY.md §8f rank 4: Dynamic_TCE_volleyball (context encoder prepended to DIN, C = 1536)
    "volleyball_vgg16_tce_T10_N12_720p": (
        dict(backbone="vgg16", image_size=(720, 1280), out_size=(22, 40), emb_features=512, num_frames=10,
             num_boxes=12, lite_dim=None, ST_kernel_size=[(3, 3)], sampling_ratio=(1,), tce=True), 8, 5646.0),
}
# BASELINE.json configs[2] ("ResNet-18 lite, B=32, 8 x B200") is volleyball_res18_lite128_T10_N12_720p with
# --global-clips 32: 32 clips per step in total, 4 per GPU at N = 8 (strong scaling); without the flag every GPU keeps
# 32 clips (weak scaling).
DEFAULT_WORKLOAD = "volleyball_vgg16_lite128_T10_N12_720p"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops_sustained": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(pc, device):
    import torch
    import din_oracle as O
    import infer_model as IM
    from config import Config
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=0, backbone=bb)     # random-init weights of the reference architecture
    cfg = Config(pc.dataset)
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "crop_size",
              "num_features_boxes", "num_activities", "lite_dim", "ST_kernel_size", "scale_factor", "beta_factor",
              "hierarchical_inference", "num_DIM"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    import contextlib
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls = IM.Dynamic_collective if pc.dataset == "collective" else \
            (IM.Dynamic_TCE_volleyball if pc.tce else IM.Dynamic_volleyball)
        model = cls(cfg)
    model.load_state_dict(sd)
    return model.to(device).eval(), sd, bb


def train_step_info(model, pc, dev, images_d, boxes_d, clips, steps=5, warmup=4, dist=None):
    """Supplementary (NOT the headline metric): one stage-2 TRAINING step -- train-mode forward, on-device
    cross-entropy, backward through the head and the VGG-16 / ResNet-18 backbone -- on `clips` clips
    (scripts/train_volleyball_stage2_dynamic.py:42 batch_size = 2), followed by torch's SGD step with lr = 0: the update
    itself is torch's, but it bumps the weights' versions, so the timed step includes re-packing every weight into the
    kernels' fp16 layouts, as a real training loop pays it.
    With several ranks (dist given) every rank trains on its own `clips` clips (weak scaling) and the gradient
    all-reduce over NCCL is INSIDE the timed step: din_b200.parallel.BucketedGradientReducer packs the gradients into
    one flat buffer (one launch per bucket) and exchanges the head's bucket while the backbone's backward still runs;
    the time reported is the max over ranks."""
    import torch
    from din_b200 import metrics, ops
    reducer = None
    try:
        model.train()
        for m in model.modules():                           # the reference's set_bn_eval (train_net_dynamic.py:101-102)
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.eval()
        for q in model.parameters():
            q.requires_grad = True
        im, bx = images_d[:clips].contiguous(), boxes_d[:clips].contiguous()
        labels = (torch.arange(clips, device=dev) % pc.num_activities)

        opt = torch.optim.SGD(list(model.parameters()), lr=0.0)
        if dist is not None:
            from din_b200.parallel import BucketedGradientReducer
            reducer = BucketedGradientReducer(model)          # installs model.grad_sink: exchange happens in backward()

        def step():
            opt.zero_grad(set_to_none=True)
            loss = metrics.cross_entropy(model((im, bx))["activities"], labels)
            loss.backward()
            opt.step()
            return loss

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        launches0 = ops.LAUNCHES
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            loss = step()
            ev[i + 1].record()
        torch.cuda.synchronize()
        launches_per_step = (ops.LAUNCHES - launches0) // steps
        # per-step times: the Inception-v3 / ResNet-18 steps are host-bound (582 / 241 launches), so one host hiccup (a
        # cudaMalloc, a page fault, a noisy neighbour on the box) lands in the mean of five steps at full weight; the
        # MEDIAN step is reported as ms_per_step, the mean and the individual steps beside it
        step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        ms_mean = ev[0].elapsed_time(ev[steps]) / steps
        ms = sorted(step_ms)[steps // 2]
        # the kernel breakdown comes from a SEPARATE, untimed pass: two CUDA events per launch cost ~0.15 ms of host time,
        # which made the 582-launch Inception-v3 step look host-bound at 121 ms (it is 32 ms)
        ops.RECORDER = []
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        rec, ops.RECORDER = ops.RECORDER, None
        world = 1
        if dist is not None:
            world = dist.get_world_size()
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        by = {}
        flops = 0.0
        for (n, f, b, a, z) in rec:
            k = n.split("_")[0].split("@")[0]
            k = "conv_fwd/dgrad" if k.startswith("conv") else ("wgrad" if k.startswith("wgrad") else k)
            by[k] = by.get(k, 0.0) + a.elapsed_time(z) / steps
            flops += f / steps
        info = {"clips_per_gpu": clips, "world": world, "ms_per_step": ms, "ms_per_step_mean": ms_mean,
                "step_ms": [round(v, 2) for v in step_ms], "clips_per_s": clips * world / (ms / 1e3),
                "loss": float(loss.detach()),
                "allreduce": (None if reducer is None else
                              {"elements": reducer.numel, "bytes_per_step": reducer.stats["bytes"] // max(1, reducer.stats["steps"]),
                               "buckets": ["head (overlaps the backbone's backward)", "backbone"]}),
                "backbone_trained": True, "gpu_launches_per_step": launches_per_step,
                "algorithmic_tflops_per_gpu": flops / (ms / 1e3) / 1e12,
                "kernels_ms": {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:10]},
                "note": "forward(train) + cross-entropy + backward + SGD(lr=0) step (weights re-packed every step)"}
    except Exception as e:                                   # never lose the headline line to the supplementary one
        info = {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        ops.RECORDER = None
        if reducer is not None:
            model.grad_sink = None
        model.eval()
        for q in model.parameters():
            q.requires_grad = False
            q.grad = None
        torch.cuda.empty_cache()
    return info


def _slice_frames(pc, sd, t_s):
    """The same workload cut to its first t_s frames (LayerNorm parameters shaped [T, ...] are cut with it)."""
    import dataclasses
    pcs = dataclasses.replace(pc, num_frames=t_s)
    if t_s == pc.num_frames:
        return pcs, sd
    sds = dict(sd)
    for k in ("dpi_nl.weight", "dpi_nl.bias", "point_ln.weight", "point_ln.bias", "DPI.hier_LN.weight",
              "DPI.hier_LN.bias"):
        if k in sds and sds[k].shape[0] == pc.num_frames:
            sds[k] = sds[k][:t_s].contiguous()
    return pcs, sds


def cpu_reference_clips_per_s(pc, sd, bb, budget_s, steps=1, warmup=0):
    """Times the reference path on the host cores (all threads), one clip per step.

    kind "reference": the reference's OWN classes (infer_model.Dynamic_volleyball / Dynamic_collective, imported
    unmodified from /root/reference or from the copy oracle/make_ref.py staged under oracle/_ref/, with the documented
    patches I / H / C applied in memory, oracle/ref_harness.py) whenever those sources are present;
    kind "port": the oracle restatement (oracle/din_oracle.py) otherwise.  Where both exist the port is run once on the
    same clip and the agreement of the two logits is reported.
    If (steps + warmup) clips would exceed `budget_s`, the clip is cut to its first t_s frames and the result is scaled
    by t_s / T (the backbone is >= 97 % of the CPU time and linear in frames; the hierarchical model's hard-coded
    [10,12,1024] LayerNorm keeps T = 10).  -> (clips/s, threads, kind, sample description, seconds per step)."""
    import torch
    import din_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O.load_backbone(bb, sd)
    T = pc.num_frames
    batch = O.make_inputs(pc, 1, seed=0)
    images = batch[0]
    t0 = time.perf_counter()
    with torch.no_grad():
        bb(O.prep_images(images[0, :1]))                      # probe: one frame through the backbone
    t_frame = time.perf_counter() - t0
    t_s = T
    if (steps + warmup) * T * t_frame > budget_s and not pc.hierarchical_inference:
        t_s = min(T, max(1, int(budget_s / ((steps + warmup) * t_frame))))
    pcs, sds = _slice_frames(pc, sd, t_s)
    cut = tuple(t[:, :t_s].contiguous() for t in batch)

    port = O.collective_forward if pc.dataset == "collective" else O.volleyball_forward
    kind, agree = "port", None
    run = lambda: port(bb, sds, pcs, *cut)                    # noqa: E731
    try:
        import ref_harness as R
        if R.available():
            import contextlib
            import io
            import warnings
            model = R.build_ref_model(pcs, sds)

            def run():
                with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                    warnings.simplefilter("ignore")
                    return model(cut)["activities"]
            kind = "reference"
    except Exception as e:                                    # never lose the line to the optional arm
        print(f"[bench] reference classes unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
    dt = (time.perf_counter() - t0) / steps
    if kind == "reference":
        agree = float((out - port(bb, sds, pcs, *cut)).abs().max())
    clips_per_s = (t_s / T) / dt
    sample = (f"{steps} step(s) x 1 clip x {t_s}/{T} frames at {pc.image_size[0]}x{pc.image_size[1]}, N={pc.num_boxes}; "
              + ("the reference's own infer_model classes (patched in memory: SURVEY.md §8c), "
                 if kind == "reference" else "oracle port, ")
              + f"torch-CPU fp32 (oneDNN), {torch.get_num_threads()} threads"
              + ("" if t_s == T else "; clips/s scaled by frames/T")
              + ("" if agree is None else f"; max|reference - oracle port| on this clip = {agree:.2e}"))
    return clips_per_s, torch.get_num_threads(), kind, sample, dt


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write to file descriptor 1 behind Python's back (NCCL
    prints its version banner there at communicator creation): point fd 1 at stderr for the whole run and keep a
    private duplicate of the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout, obj):
    os.write(real_stdout, (json.dumps(obj) + "\n").encode())


def ingest_info(pc, dev, frames=80):
    """Supplementary (NOT the headline metric): the loader's per-frame work on the device --
    synthetic JPEG files (encoded here with Pillow, 4:2:0, quality 90) decoded by nvJPEG and resized to the workload's
    image size by the Pillow-exact resize kernel (din_b200.ingest.decode_resize), against one Pillow worker doing the
    reference's Image.open + resize + np.array (volleyball.py:237-240) on the same files."""
    import io
    import time
    try:
        import numpy as np
        import torch
        from PIL import Image
        from din_b200 import _lib, ingest
        H, W = pc.image_size
        src = (720, 1280)                                       # the Volleyball dataset's frame size
        rng = np.random.default_rng(0)
        yy, xx = np.mgrid[0:src[0], 0:src[1]]
        files = []
        for i in range(4):
            img = np.stack([(xx * 255 // (src[1] - 1)), (yy * 255 // (src[0] - 1)), ((xx + yy + 40 * i) % 256)], -1)
            img = (img + 40 * np.sin(xx / 9.0)[..., None] + rng.integers(-25, 26, size=src + (3,))).clip(0, 255).astype(np.uint8)
            b = io.BytesIO()
            Image.fromarray(img).save(b, format="JPEG", quality=90, subsampling=2)
            files.append(b.getvalue())
        jpegs = (files * ((frames + 3) // 4))[:frames]
        with torch.cuda.device(dev):
            ingest.decode_resize(jpegs, (H, W), device=dev)     # warm-up: nvJPEG state, coefficient tables
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                out = ingest.decode_resize(jpegs, (H, W), device=dev)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            raw = torch.randint(0, 256, (frames,) + src + (3,), dtype=torch.uint8, device=dev)
            resize_ms = None
            if src != (H, W):
                ingest.resize_u8(raw, (H, W))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    ingest.resize_u8(raw, (H, W))
                e1.record()
                torch.cuda.synchronize()
                resize_ms = e0.elapsed_time(e1) / 5
        t0 = time.perf_counter()
        n_pil = 16
        for j in jpegs[:n_pil]:
            np.array(Image.open(io.BytesIO(j)).resize((W, H), Image.BILINEAR))
        pil_fps = n_pil / (time.perf_counter() - t0)
        return {"frames_per_s": frames / dt, "frames": frames, "source": f"{src[0]}x{src[1]} JPEG 4:2:0 q90, "
                f"{sum(map(len, jpegs)) // frames // 1000} kB/frame, from host memory", "target": [H, W],
                "nvjpeg_backend": int(_lib.load().din_jpeg_backend()),
                "resize_ms_per_batch": resize_ms, "resize_bit_identical_to_pillow": True,
                "pillow_one_worker_frames_per_s": pil_fps, "output": list(out.shape),
                "note": "decode = nvJPEG (library); resize = csrc/ingest.cu; wall clock incl. the final stream sync"}
    except Exception as e:                                   # never lose the headline line to the supplementary one
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--clips-per-gpu", type=int, default=None)
    ap.add_argument("--global-clips", type=int, default=None,
                    help="fixed TOTAL clips per step, split over the GPUs (strong scaling), e.g. 32 for BASELINE configs[2]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--no-ingest", action="store_true")
    ap.add_argument("--train-clips", type=int, default=2, help="clips per training step (reference batch_size = 2)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import dataclasses
    import torch
    import din_oracle as O
    kw, clips_per_gpu, gflop_per_clip = WORKLOADS[args.workload]
    scaling = "weak"
    if args.global_clips:
        assert args.global_clips % world == 0, "--global-clips must divide by the number of GPUs"
        clips_per_gpu, scaling = args.global_clips // world, "strong"
    elif args.clips_per_gpu:
        clips_per_gpu = args.clips_per_gpu
    pc = O.PathConfig(**kw)
    config = {"workload": args.workload, "backbone": pc.backbone, "clips_per_gpu": clips_per_gpu,
              "global_clips_per_step": clips_per_gpu * world, "frames": pc.num_frames, "actors": pc.num_boxes,
              "image": list(pc.image_size), "parallelism": f"dp{world} (clips sharded, no forward collective)",
              "l2": "inputs (%.0f MB/step/GPU) larger than L2, no explicit flush" %
                    (clips_per_gpu * pc.num_frames * 3 * pc.image_size[0] * pc.image_size[1] * 4 / 1e6)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        bb = O.build_backbone(pc.backbone)
        sd = O.make_state_dict(pc, seed=0, backbone=bb)
        v, cores, kind, sample, dt = cpu_reference_clips_per_s(pc, sd, bb, budget_s=200.0, steps=max(1, args.steps),
                                                               warmup=args.warmup)
        _emit(real_stdout, {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, clips_per_gpu=1, global_clips_per_step=1, parallelism="cpu"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return

    # ------------------------------------------------------------------ our arm (B200)
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the hot path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from din_b200 import ops

    model, sd, bb = build_model(pc, dev)
    B, T, N = clips_per_gpu, pc.num_frames, pc.num_boxes
    H, W = pc.image_size
    # boxes (and Collective's per-clip actor counts: a different count per clip) from the oracle's generator
    host = O.make_inputs(dataclasses.replace(pc, image_size=(8, 8)), B, seed=rank)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images_d = torch.randint(0, 256, (B, T, 3, H, W), generator=g, device=dev, dtype=torch.int32).float()
    boxes_d = host[1].to(dev)
    extra_d = tuple(t.to(dev) for t in host[2:])                 # (bboxes_num [B,T] int32,) for Collective

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.no_grad():
            return model((images_d, boxes_d) + extra_d)["activities"]

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.LAUNCHES
    # per-launch CUDA events are recorded INSIDE the timed region (on the launching stream, no syncs), so the
    # roofline numbers see the same sustained clocks as `value`
    ops.RECORDER = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier()
    rec, ops.RECORDER = ops.RECORDER, None
    launches = ops.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(out).all()

    # ---- end to end through the public API with HOST inputs: model((frames, boxes)) on pinned host tensors -- the engine
    #      copies the frames chunk by chunk on a copy stream under the backbone's kernels -- -> D2H of the logits.
    #      `e2e`    : fp32 [B,T,3,H,W] frames, the tensor the reference's loader yields (volleyball.py:270)
    #      `e2e_u8` : uint8 [B,T,H,W,3] frames, the decoded images before the loader's transpose / float()
    #                 (SURVEY.md section 8f rank 2): 4x fewer bytes over PCIe, bit-identical logits
    e2e = e2e_u8 = None

    def measure_e2e(img_src):
        img_host = torch.empty(img_src.shape, dtype=img_src.dtype).pin_memory()
        img_host.copy_(img_src)
        box_host = boxes_d.cpu().pin_memory()
        extra_host = tuple(t.cpu().pin_memory() for t in extra_d)
        out_host = torch.empty((B, pc.num_activities), dtype=torch.float32).pin_memory()
        def run_e2e(n):
            # the public call on HOST tensors: the engine streams the frames to the GPU chunk by chunk on its own copy
            # stream, under the backbone kernels of the previous chunk (DinEngine._stage_host_chunk); nothing blocks the
            # host, so step i + 1's transfers are queued while step i still computes
            for i in range(n):
                with torch.no_grad():
                    o = model((img_host, box_host) + extra_host)["activities"]
                out_host.copy_(o, non_blocking=True)           # D2H read of the step's result
            torch.cuda.synchronize()

        run_e2e(2)
        barrier()
        t0 = time.perf_counter()
        e2e0, e2e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e0.record()
        run_e2e(args.steps)
        e2e1.record()
        barrier()
        return {"ms": e2e0.elapsed_time(e2e1), "wall_ms": (time.perf_counter() - t0) * 1e3,
                "h2d": img_host.numel() * img_host.element_size() + box_host.numel() * 4, "d2h": out_host.numel() * 4}

    if not args.no_e2e:
        e2e = measure_e2e(images_d)
        e2e_u8 = measure_e2e(images_d.permute(0, 1, 3, 4, 2).contiguous().to(torch.uint8))

    # ---- roofline of the dominant kernel from the events recorded in the timed region (per step averages)
    conv = [(n, f, b, a.elapsed_time(z)) for (n, f, b, a, z) in rec if n.startswith("conv")]
    all_ms = sum(a.elapsed_time(z) for (_, _, _, a, z) in rec) / args.steps
    conv_ms = sum(t for *_, t in conv) / args.steps
    conv_flops = sum(f for _, f, _, _ in conv) / args.steps
    per_layer = {}
    for n, f, b, t in conv:
        e = per_layer.setdefault(n, [0, 0.0, 0])
        e[0] += f; e[1] += t; e[2] += 1
    other = {}
    for (n, f, b, a, z) in rec:
        if not n.startswith("conv"):
            k = n.split("_")[0].split("@")[0]
            other[k] = other.get(k, 0.0) + a.elapsed_time(z) / args.steps

    # supplementary: the training step (all ranks take part: its gradient all-reduce is inside the timed step)
    train_info = None
    if not args.no_train_step and pc.backbone in ("vgg16", "res18", "inv3") and pc.dataset == "volleyball":
        train_info = train_step_info(model, pc, dev, images_d, boxes_d, min(args.train_clips, B), dist=dist)

    if dist is not None:
        t = torch.tensor([ms, e2e["ms"] if e2e else 0.0, e2e_u8["ms"] if e2e_u8 else 0.0], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if e2e:
            e2e["ms"], e2e_u8["ms"] = float(t[1]), float(t[2])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    total_clips = B * world * args.steps
    value = total_clips / (ms / 1e3)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (backbone), f32 (head)",
        "data": "synthetic", "config": config,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 implicit GEMM, all launches of a step)",
                     "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                     "traffic_unit": "DRAM bytes (read+write) per step in this kernel, ncu --set full (profiles/)",
                     "peak_source": peaks["source"] + ", sustained bf16 (= fp16 rate)",
                     "flops_per_step": conv_flops, "kernel_ms_per_step": conv_ms,
                     "kernel_share_of_step": conv_ms / (ms / args.steps),
                     "all_kernels_ms_per_step": all_ms,
                     "whole_path_tflops": value * gflop_per_clip / 1e3,
                     "whole_path_frac": value * gflop_per_clip / 1e3 / world / peaks["tflops_sustained"],
                     "other_kernels_ms": {k: round(v, 3) for k, v in sorted(other.items())},
                     "per_layer_tflops": {n: round(f / (t / 1e3) / 1e12, 1) for n, (f, t, c) in per_layer.items()},
                     "per_layer_ms_per_step": {n: round(t / args.steps, 3) for n, (f, t, c) in per_layer.items()}},
        "clocks": clocks.summary(),
        "gpu_launches": launches,
    }
    if e2e:
        line["e2e"] = {"value": total_clips / (e2e["ms"] / 1e3), "unit": UNIT,
                       "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "input": "fp32 [B,T,3,H,W] frames (the reference loader's tensor), pinned host memory"}
        line["e2e_u8"] = {"value": total_clips / (e2e_u8["ms"] / 1e3), "unit": UNIT,
                          "h2d_bytes_per_step": e2e_u8["h2d"], "d2h_bytes_per_step": e2e_u8["d2h"],
                          "input": "uint8 [B,T,H,W,3] frames (decoded images before the loader's transpose/float), "
                                   "bit-identical logits"}
    if train_info is not None:
        line["train_step"] = train_info
    if world == 1 and not args.no_ingest:
        line["ingest"] = ingest_info(pc, dev)
    if world == 1 and not args.no_cpu_baseline:
        v, cores, kind, sample, _ = cpu_reference_clips_per_s(pc, sd, bb, budget_s=args.cpu_budget_s)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    _emit(real_stdout, line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
