"""cProfile of the host side of a training step (which Python / torch calls keep the GPU waiting?)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
name = os.environ.get("W", "volleyball_inv3_full_T10_N12_720p")
pc_kw, B, _ = bench.WORKLOADS[name]
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import din_oracle as O
from din_b200 import metrics
pc = O.PathConfig(**pc_kw)
dev = torch.device("cuda:0")
model, sd, bb = bench.build_model(pc, dev)
images, boxes = O.make_inputs(pc, 2, seed=0)[:2]
im, bx = images.to(dev), boxes.to(dev)
model.train()
for m in model.modules():
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.eval()
for q in model.parameters():
    q.requires_grad = True
labels = torch.arange(2, device=dev) % pc.num_activities
opt = torch.optim.SGD(list(model.parameters()), lr=0.0)


def step():
    opt.zero_grad(set_to_none=True)
    loss = metrics.cross_entropy(model((im, bx))["activities"], labels)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"3 steps: host {1e3 * (t1 - t0) / 3:.1f} ms/step, with sync {1e3 * (t2 - t0) / 3:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
