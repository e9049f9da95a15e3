"""Where does the HOST time of a training step go?  cProfile over a few stage-2 training steps (2 clips, backbone
trained) -- dev tool, prints the top functions by cumulative time.  usage: train_host_profile.py [vgg16|res18]"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"))
import bench  # noqa: E402
import din_oracle as O  # noqa: E402
from din_b200 import metrics  # noqa: E402

bbname = sys.argv[1] if len(sys.argv) > 1 else "res18"
name = f"volleyball_{bbname}_lite128_T10_N12_720p"
kw, _, _ = bench.WORKLOADS[name]
pc = O.PathConfig(**kw)
dev = torch.device("cuda:0")
model, sd, bb = bench.build_model(pc, dev)
images, boxes = (t.to(dev) for t in O.make_inputs(pc, 2, seed=0))
model.train()
BN_BATCH = len(sys.argv) > 2 and sys.argv[2] == "bn"      # BatchNorm on batch statistics (no cfg.set_bn_eval)
for m in model.modules():
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and not BN_BATCH:
        m.eval()
for q in model.parameters():
    q.requires_grad = True
labels = torch.arange(2, device=dev) % pc.num_activities
opt = torch.optim.SGD(list(model.parameters()), lr=0.0)


def step():
    opt.zero_grad(set_to_none=True)
    loss = metrics.cross_entropy(model((images, boxes))["activities"], labels)
    loss.backward()
    opt.step()
    return loss


if os.environ.get("DIN_NCU") == "1":             # under ncu: one warm-up step, one profiled step, nothing else
    step()
    torch.cuda.synchronize()
    step()
    torch.cuda.synchronize()
    sys.exit(0)
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t_host = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5
print(f"{bbname}{' (BatchNorm on batch statistics)' if BN_BATCH else ''}: host issue time {t_host * 1e3:.1f} ms / step, with final sync {t_all * 1e3:.1f} ms / step")
if os.environ.get("DIN_KINETO", "1") == "1":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
