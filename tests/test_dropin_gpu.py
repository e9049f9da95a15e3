"""Drop-in level of SURVEY.md §4: the reference's unmodified `scripts/train_volleyball_stage2_dynamic.py` ->
`train_net_dynamic.train_net` runs on this package's modules on a real B200 (tests/tools/dropin_run.py), and the
`nn.DataParallel` wrap the trainer applies (train_net_dynamic.py:96) does not rebuild the plan per call."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in ("/root/reference", os.path.join(ROOT, "oracle", "_ref", "reference"))
            if os.path.exists(os.path.join(p, "train_net_dynamic.py"))), None)


@pytest.mark.skipif(REF is None, reason="reference sources not staged (python oracle/make_ref.py)")
@pytest.mark.parametrize("script,model", [("train_volleyball_stage2_dynamic.py", "Dynamic_volleyball"),
                                          ("train_volleyball_stage2_dynamic_tce.py", "Dynamic_TCE_volleyball")])
def test_unmodified_reference_script_trains_and_checkpoints(tmp_path, cuda, script, model):
    """2 training steps (batch 2, T = 10, 720p, VGG-16 trained, nn.DataParallel wrap) + 1 test step + checkpoint,
    then the checkpoint re-loaded into a fresh drop-in model; the same for the TCE sibling's script."""
    env = dict(os.environ, DIN_OFFLINE="1")
    env.pop("CUDA_VISIBLE_DEVICES", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "dropin_run.py"), "--ref", REF,
                          "--workdir", str(tmp_path), "--epochs", "1", "--devices", "0", "--script", script],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    info = json.loads(out.stdout.strip().splitlines()[-1])
    print(info)
    assert len(info["losses"]) == 2, info                       # one 'Train' and one 'Test' epoch line
    assert all(0.0 < v < 50.0 for v in info["losses"]), info    # finite
    assert info["checkpoint"].startswith("stage2_epoch1_") and info["epochs"] == 1
    assert info["reloaded_logits_finite"] and info["data_parallel"] and info["model"] == model
    assert info["optimizer_state_tensors"] > 30                 # Adam state for the backbone + head parameters


def _small_model(dev):
    import din_oracle as O
    import infer_model as IM
    from config import Config
    pc = O.PathConfig(backbone="vgg16", image_size=(64, 96), out_size=(2, 3), num_frames=2, num_boxes=3)
    bb = O.build_backbone("vgg16")
    sd = O.make_state_dict(pc, seed=0, backbone=bb)
    cfg = Config("volleyball")
    cfg.log_path = None
    for k in ("backbone", "image_size", "out_size", "emb_features", "num_frames", "num_boxes", "lite_dim",
              "ST_kernel_size", "beta_factor"):
        setattr(cfg, k, getattr(pc, k))
    cfg.sampling_ratio = list(pc.sampling_ratio)
    model = IM.Dynamic_volleyball(cfg)
    model.load_state_dict(sd)
    images, boxes = O.make_inputs(pc, 4, seed=0)
    return model.to(dev), images.to(dev), boxes.to(dev)


def test_data_parallel_wrap_keeps_the_plan(cuda):
    model, images, boxes = _small_model(cuda)
    model.eval()
    with torch.no_grad():
        want = model((images, boxes))["activities"]
    assert model._plans.builds == 1
    dp = torch.nn.DataParallel(model, device_ids=[0])
    with torch.no_grad():
        for _ in range(3):
            got = dp((images, boxes))["activities"]
    assert model._plans.builds == 1                             # not rebuilt by the wrap or by repeated calls
    assert torch.equal(got, want)
    # a training step through the wrap: gradients arrive on the wrapped model's parameters, and only the optimizer
    # step invalidates the plan
    dp.train()
    opt = torch.optim.SGD([p for p in dp.parameters() if p.requires_grad], lr=1e-3)
    loss = torch.nn.functional.cross_entropy(dp((images, boxes))["activities"], torch.tensor([0, 1, 2, 3], device=cuda))
    loss.backward()
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)
    builds = model._plans.builds
    dp.eval()
    with torch.no_grad():
        dp((images, boxes))
    assert model._plans.builds == builds
    opt.step()
    with torch.no_grad():
        dp((images, boxes))
    assert model._plans.builds == builds + 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_two_devices_one_plan_per_device(cuda):
    model, images, boxes = _small_model(cuda)
    model.eval()
    with torch.no_grad():
        want = model((images, boxes))["activities"]
    dp = torch.nn.DataParallel(model, device_ids=[0, 1])
    with torch.no_grad():
        for _ in range(3):
            got = dp((images, boxes))["activities"]
    assert model._plans.builds == 2, model._plans.builds        # cuda:0 (reused from above) + cuda:1, once each
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    dp.train()
    loss = torch.nn.functional.cross_entropy(dp((images, boxes))["activities"], torch.tensor([0, 1, 2, 3], device=cuda))
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)


def test_logits_do_not_depend_on_batch_composition(cuda):
    """A clip's logits are the same whether it runs in a batch of 4 or of 2 (what nn.DataParallel's scatter does to it; the
    concurrent-replica side of that is test_data_parallel_two_devices_one_plan_per_device and, for the per-thread
    inference flag, tests/test_dropin_cpu.py::test_inference_flag_is_per_thread)."""
    model, images, boxes = _small_model(cuda)
    model.eval()
    with torch.no_grad():
        want = model((images, boxes))["activities"]
        halves = [model((images[i:i + 2].contiguous(), boxes[i:i + 2].contiguous()))["activities"] for i in (0, 2)]
    assert torch.equal(torch.cat(halves), want)
