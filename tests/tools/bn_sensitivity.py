"""How far do the ORACLE's own gradients move under an fp16-sized perturbation (every conv weight x (1 + 3e-4 N(0,1)))?
CPU only.  Result (res18_lite, 2 clips x 3 frames at 96x160), committed in profiles/README.md item 13:
  BatchNorm eval mode   : logits 2.5e-3, backbone gradients worst 6.7e-2 (median 3.8e-2), head 1.4e-2
  BatchNorm batch stats : logits 5.9e-3, backbone gradients worst 1.9e-1 (median 1.2e-1), head 1.3e-1
i.e. the batch-statistics network is ~3x more sensitive; the CUDA path's deviations from the oracle (6.8e-2 / 1.8e-1) are
of exactly this size.  Sets the tolerances of tests/test_backward_gpu.py."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import din_oracle as O
from test_oracle_cpu import _pc_from
fx = torch.load(os.path.join(ROOT, 'tests', 'golden', 'bntrain_res18_lite.pt'))
pc = _pc_from(fx["config"])
def run(bn_train, noise):
    bb = O.build_backbone(pc.backbone)
    sd = O.make_state_dict(pc, seed=0, backbone=bb)
    O.load_backbone(bb, sd); bb.train(bn_train)
    batch = list(O.make_inputs(pc, fx["B"], seed=0))
    if noise:
        g = torch.Generator().manual_seed(1)
        # fp16-like relative perturbation of every conv weight (the activations' rounding is harder to emulate)
        for n, q in bb.named_parameters():
            if q.dim() == 4:
                q.data.mul_(1 + noise * torch.randn(q.shape, generator=g))
    return O.head_grads(bb, sd, pc, fx["labels"], *batch, train_backbone=True)
for bn_train in (False, True):
    l0, loss0, g0 = run(bn_train, 0)
    l1, loss1, g1 = run(bn_train, 3e-4)
    rel = {k: float((g1[k]-g0[k]).norm()/g0[k].norm()) for k in g0}
    bbk = [k for k in rel if k.startswith('backbone.')]
    print("bn_train", bn_train, "loss", float(loss0), float(loss1), "logits d", float((l1-l0).abs().max()),
          "worst backbone", max(rel[k] for k in bbk), "median", sorted(rel[k] for k in bbk)[len(bbk)//2],
          "worst head", max(rel[k] for k in rel if not k.startswith('backbone.')))
