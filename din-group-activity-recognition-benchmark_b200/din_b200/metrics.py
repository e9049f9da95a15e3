"""Loss and epoch metrics kept on the device.

The reference's loops compute, per step, F.cross_entropy, argmax, a correct count pulled with `.item()`,
`loss.item()` for the AverageMeter and ConfusionMeter.add(...) which copies both tensors to the host
(train_net_dynamic.py:191-199,217, 258-292; utils.py:193-264): three device->host synchronisations per
step.  Here one kernel launch (din_ce_metrics_f32) produces the loss, d(loss)/d(logits), the correct count,
and accumulates the confusion matrix and the meter sums in device memory; `DeviceMeters.value()` reads
them once, when the epoch summary is printed.
"""
import numpy as np
import torch

from . import ops


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_weight, loss_scale, conf, meters):
        need = ctx.needs_input_grad[0]
        loss, correct, dlogits = ops.ce_metrics(logits.detach().contiguous(), labels, class_weight=class_weight,
                                                loss_scale=loss_scale, conf=conf, meters=meters, want_grad=need)
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(correct)
        return loss.view(()), correct

    @staticmethod
    def backward(ctx, grad_loss, _grad_correct):
        (dlogits,) = ctx.saved_tensors
        # the kernel already wrote d(loss)/d(logits); only the upstream scalar (1 for `loss.backward()`) is applied
        return dlogits * grad_loss, None, None, None, None, None


def cross_entropy(logits, labels, weight=None, loss_scale=1.0, meters=None):
    """Drop-in for `F.cross_entropy(scores, labels, weight=...) * loss_scale` (train_net_dynamic.py:193,204)
    on CUDA fp32 logits; differentiable w.r.t. `logits`.  With `meters` (a DeviceMeters) the same launch
    also updates the accuracy / confusion / loss meters."""
    labels = labels.to(torch.int64).contiguous()
    conf = meters.conf if meters is not None else None
    acc = meters.sums if meters is not None else None
    loss, _ = _CrossEntropyFn.apply(logits, labels, weight, float(loss_scale), conf, acc)
    return loss


class DeviceMeters:
    """AverageMeter (loss), AverageMeter (accuracy) and ConfusionMeter (utils.py:193-276) as device buffers."""

    def __init__(self, num_classes, device):
        self.k = num_classes
        self.conf = torch.zeros((num_classes, num_classes), dtype=torch.int32, device=device)
        self.sums = torch.zeros((4,), dtype=torch.float64, device=device)   # sum(loss*b), sum(b), sum(correct), steps

    def reset(self):
        self.conf.zero_()
        self.sums.zero_()

    def update(self, logits, labels, weight=None, loss_scale=1.0):
        """Metrics only (evaluation loop): returns the step's loss tensor (device, no sync)."""
        with torch.no_grad():
            return cross_entropy(logits, labels, weight=weight, loss_scale=loss_scale, meters=self)

    def value(self):
        """One device->host read.  Keys follow the reference's epoch summary (train_net_dynamic.py:226-233)."""
        conf = self.conf.cpu().numpy()
        s = self.sums.cpu().numpy()
        n = max(float(s[1]), 1.0)
        class_sum = conf.sum(axis=1).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            per_class = np.diag(conf).astype(np.float32) / class_sum          # MPCA (utils.py:279-289)
        return {"loss": float(s[0]) / n, "activities_acc": float(s[2]) / n * 100.0, "activities_conf": conf,
                "activities_MPCA": float(np.mean(per_class) * 100.0), "samples": int(s[1]), "steps": int(s[3])}
