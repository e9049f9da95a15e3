"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: clip sharding, logits gather, max-over-ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions_exactly():
    from din_b200.parallel import shard_range
    for n in (0, 1, 2, 7, 8, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            lens = [b - a for a, b in spans]
            assert max(lens) - min(lens) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _FakeModel:
    """Stands in for Dynamic_volleyball on CPU: per-clip logits that depend only on that clip."""
    class cfg:
        num_activities = 8

    def __call__(self, batch):
        images, boxes = batch
        base = images.flatten(1).mean(1, keepdim=True) + boxes.flatten(1).sum(1, keepdim=True)
        return {"activities": base * torch.arange(1, 9, dtype=torch.float32)}


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from din_b200.parallel import max_over_ranks, sharded_forward
    g = torch.Generator().manual_seed(0)
    images = torch.rand(n_clips, 2, 3, 4, 5, generator=g)
    boxes = torch.rand(n_clips, 2, 3, 4, generator=g)
    out = sharded_forward(_FakeModel(), (images, boxes))
    ref = _FakeModel()((images, boxes))["activities"]
    ok = torch.allclose(out, ref) and out.shape == ref.shape
    slowest = max_over_ranks(10.0 + rank)
    q.put((rank, bool(ok), slowest))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [1, 5, 8])
def test_sharded_forward_world2_gloo(n_clips):
    import sys
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "din-group-activity-recognition-benchmark_b200")
    os.environ["PYTHONPATH"] = pkg + os.pathsep + os.environ.get("PYTHONPATH", "")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]      # identical, correctly ordered logits on both ranks
    assert [r[2] for r in res] == [11.0, 11.0]      # max over ranks


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from din_b200.parallel import GradientAllReducer
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    extra = torch.nn.Parameter(torch.zeros(4))            # never used: no gradient on any rank
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    x = torch.full((2, 5), float(rank + 1))
    lin(x).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    reducer = GradientAllReducer(list(lin.parameters()) + [extra, frozen])
    n = reducer()
    # the mean over ranks of grads computed with x = 1 and x = 2: weight grad rows = 2 * mean(1, 2) = 3, bias = 2
    ok = (n == 15 + 3 + 4 and torch.allclose(lin.weight.grad, torch.full((3, 5), 3.0))
          and torch.allclose(lin.bias.grad, torch.full((3,), 2.0)) and torch.equal(extra.grad, torch.zeros(4))
          and frozen.grad is None and not torch.equal(local[0], lin.weight.grad))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    """One flat all-reduce per step averages every parameter gradient over the ranks."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "din-group-activity-recognition-benchmark_b200")
    os.environ["PYTHONPATH"] = pkg + os.pathsep + os.environ.get("PYTHONPATH", "")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from din_b200.parallel import BucketedGradientReducer

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Linear(5, 3)
            self.fc_emb_1 = torch.nn.Linear(3, 7)           # 21 + 7 elements: slots are padded to multiples of 4
            self.unused = torch.nn.Parameter(torch.zeros(6))
            self.frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)

    torch.manual_seed(0)
    net = Net()
    reducer = BucketedGradientReducer(net)
    assert net.grad_sink is reducer and reducer.active()
    ok = True
    for step in range(2):                                   # the flat buffer is re-used across steps
        # rank r holds r + 1 of the 3 clips of the global batch; its "mean over local clips" gradient is (r + 1 + step)
        local, total = rank + 1, 3
        val = float(rank + 1 + step)
        grads = {n: torch.full_like(p, val) for n, p in net.named_parameters() if n not in ("unused", "frozen")}
        for p in net.parameters():
            p.grad = None
        reducer.set_batch(local, total)
        reducer("head", grads)                              # issued while the "backbone backward" would still run
        reducer("backbone", grads)
        reducer.finish()
        want = (1 / 3) * (1 + step) + (2 / 3) * (2 + step)  # sum_r (n_r / n) g_r = the global-batch mean gradient
        for n, p in net.named_parameters():
            if n == "frozen":
                ok &= p.grad is None
            elif n == "unused":
                ok &= bool(torch.equal(p.grad, torch.zeros(6)))
            else:
                ok &= bool(torch.allclose(p.grad, torch.full_like(p, want), atol=1e-6))
                ok &= p.grad.data_ptr() % 16 == 0 and p.grad.is_contiguous()
        # .grad tensors are views of ONE flat buffer
        base = reducer._flat.data_ptr()
        ok &= all(base <= p.grad.data_ptr() < base + 4 * reducer.numel for p in net.parameters() if p.grad is not None)
    q.put((rank, bool(ok), reducer.stats["steps"]))
    dist.destroy_process_group()


def test_bucketed_gradient_reducer_world2_gloo():
    """Two buckets (head first, backbone last), each all-reduced asynchronously as soon as it is final; unequal shards
    are weighted by local_clips / global_clips so that the result is the global-batch mean gradient."""
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "din-group-activity-recognition-benchmark_b200")
    os.environ["PYTHONPATH"] = pkg + os.pathsep + os.environ.get("PYTHONPATH", "")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:] for r in res] == [(True, 2), (True, 2)]


def test_gradient_allreduce_weights_unequal_shards():
    from din_b200.parallel import _rank_weight
    assert _rank_weight(1, 4) == 0.25 and _rank_weight(0, 4) == 0.0
    with pytest.raises(ValueError):
        _rank_weight(5, 4)
