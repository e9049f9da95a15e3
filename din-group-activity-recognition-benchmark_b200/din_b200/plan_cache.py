"""Which tensors a drop-in model's plan is built from, and when it has to be rebuilt.

A plan (packed fp16 weights, tensor maps, workspaces) is cached per (model, device) and keyed on the identity and
`_version` of the OWNER model's tensors: an optimizer step or `load_state_dict` bumps the versions and the next
forward re-packs; nothing else does.

`nn.DataParallel` (reference train_net_dynamic.py:96) needs care.  With one visible device it calls the wrapped
module directly, so nothing changes.  With several, every forward works on fresh *replicas*
(torch/nn/parallel/replicate.py): shallow copies of the module whose parameters are re-broadcast tensors, kept as
plain attributes / `_former_parameters` (`named_parameters()` and `state_dict()` of a replica list no parameters at
all).  Keying on a replica's tensors would rebuild the whole plan on every step on every device.  Instead each model
keeps `_owner = [self]`: the shallow copy hands that same list to every replica, so a replica finds the original
module, keys on ITS tensor versions and parks its plan in the owner's per-device table.  A frozen model in
evaluation mode therefore packs once per device, and a training step re-packs exactly when the optimizer stepped.
"""
import collections

import torch


def is_replica(module):
    return bool(getattr(module, "_is_replica", False))


def named_tensors(module, buffers=True):
    """state_dict()-ordered {name: tensor} that also works on a DataParallel replica."""
    if not is_replica(module):
        sd = module.state_dict() if buffers else collections.OrderedDict(module.named_parameters())
        return collections.OrderedDict(sd)
    out = collections.OrderedDict()
    for prefix, m in module.named_modules():
        dot = prefix + "." if prefix else ""
        for k, v in getattr(m, "_former_parameters", {}).items():
            if v is not None:
                out[dot + k] = v
        if buffers:
            for k, v in m._buffers.items():
                if v is not None and k not in m._non_persistent_buffers_set:
                    out[dot + k] = v
    return out


def trainable(module):
    """[(name, tensor)] of everything that wants a gradient (parameters; on a replica the broadcast copies, through
    which autograd reaches the owner's parameters)."""
    return [(n, p) for n, p in named_tensors(module, buffers=False).items() if p.requires_grad]


def owner_of(module):
    ref = getattr(module, "_owner", None)
    return ref[0] if ref else module


def version_key(module, buffers=True, prefix=""):
    """Identity + version of the OWNER's tensors (optionally one sub-tree / parameters only)."""
    own = owner_of(module)
    sub = own
    for part in filter(None, prefix.split(".")):
        sub = getattr(sub, part)
    tensors = sub.state_dict().values() if buffers else sub.parameters()
    return tuple((t.data_ptr(), t._version) for t in tensors)


def device_of(module):
    for t in named_tensors(module).values():
        return t.device
    raise RuntimeError("model without tensors")


def require_cuda(dev):
    if dev.type != "cuda":
        raise RuntimeError("the DIN hot path runs on sm_100a only: move the model to a CUDA device "
                           "(there is no CPU fallback)")


class PlanTable:
    """Per-device plan slots living on the owner model (shared with its replicas by reference)."""

    def __init__(self):
        self.slots = {}
        self.builds = 0          # how many times a plan was (re)built: tests assert on it

    def get(self, dev):
        return self.slots.get(str(dev))

    def put(self, dev, **slot):
        self.builds += 1
        self.slots[str(dev)] = slot
        return slot

    # a model is deep-copied / pickled with its plans dropped (they hold raw device pointers)
    def __deepcopy__(self, memo):
        return PlanTable()

    def __reduce__(self):
        return (PlanTable, ())
