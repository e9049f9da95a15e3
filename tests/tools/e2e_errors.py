"""Print end-to-end logits errors (CUDA path vs CPU oracle) for a list of configurations / seeds."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import din_oracle as O  # noqa: E402
from test_e2e_gpu import _pc, _run_case  # noqa: E402

dev = torch.device("cuda:0")
cases = []
for seed in (0, 1, 2):
    cases.append(("inv3 139x203", _pc("inv3", (139, 203), emb_features=1056, num_frames=2, num_boxes=4, lite_dim=None), 2, seed))
cases.append(("inv3 720p T=2", _pc("inv3", (720, 1280), emb_features=1056, num_frames=2, num_boxes=12, lite_dim=None), 1, 0))
for seed in (1, 2):
    cases.append(("vgg16 96x160", _pc("vgg16", (96, 160), num_frames=3, num_boxes=4), 2, seed))
    cases.append(("res18 96x160", _pc("res18", (96, 160), num_frames=3, num_boxes=4), 2, seed))
    cases.append(("collective res18", _pc("res18", (96, 144), dataset="collective", num_frames=3, num_boxes=13, lite_dim=None, ST_kernel_size=(3, 3), num_activities=4), 3, seed))
for name, pc, B, seed in cases:
    t0 = time.time()
    try:
        out, ref = _run_case(dev, pc, B, seed=seed)
    except AssertionError as e:
        print("   ^ over tolerance:", str(e)[:80])
    print(f"   ({name}, seed {seed}, {time.time() - t0:.1f}s)")
