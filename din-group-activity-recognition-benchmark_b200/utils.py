"""Host-side helpers under the reference's `utils` module name (reference utils.py:8-300).

The drop-in recipe puts this package BEFORE the reference on `sys.path`, so this file is
what `from utils import *` resolves to in the reference's unmodified trainers (`train_net_dynamic.py:15`,
`train_net.py`).  It therefore provides every name those trainers take from `utils`: `show_config` (:35),
`show_epoch_info` (:126,131), `AverageMeter` / `Timer` / `ConfusionMeter` (:161-165), `MPCA` (:232),
`print_log`, `log_final_exp_result`, the pairwise-distance helpers, `MAC2FLOP`, and -- like the reference's
star-import -- the module names `torch`, `np`, `nn`, `time`, `transforms`.  All of it is host bookkeeping
(SURVEY.md §2 #15); nothing here is on the device path.  `prep_images` is kept for API compatibility only: on
the hot path it is fused into the stem convolution kernel and the models of this package never call it.

Names this file does not define (`MADmeter`, `sincos_encoding_2d`: no trainer uses them) are taken from the
reference's own `utils.py` when one sits further down `sys.path`.
"""
import os
import pickle
import sys
import time

import numpy as np
import torch
import torch.nn as nn

try:                                                       # the reference's star-import leaks this name too
    import torchvision.transforms as transforms
except Exception:                                          # pragma: no cover
    transforms = None


def prep_images(images):
    """(x / 255 - 0.5) * 2  (reference utils.py:8-19)."""
    return images.div(255.0).sub(0.5).mul(2.0)


def calc_pairwise_distance(X, Y):
    """Euclidean distances between the rows of X [N,D] and Y [M,D] -> [N,M] (reference utils.py:42-54)."""
    sq = X.pow(2).sum(dim=1)[:, None] - 2.0 * (X @ Y.t()) + Y.pow(2).sum(dim=1)[None, :]
    return sq.sqrt()


def calc_pairwise_distance_3d(X, Y):
    """Batched form: X [B,N,D], Y [B,M,D] -> [B,N,M] (reference utils.py:56-73)."""
    sq = X.pow(2).sum(dim=2)[:, :, None] - 2.0 * torch.bmm(X, Y.transpose(1, 2)) + Y.pow(2).sum(dim=2)[:, None, :]
    return sq.sqrt()


def print_log(file_path, *args):
    """print, and append the same line to `file_path` unless it is None (reference utils.py:101-105)."""
    print(*args)
    if file_path is not None:
        with open(file_path, "a") as f:
            print(*args, file=f)


def show_config(cfg):
    """Dump every attribute of the config object to the log (reference utils.py:107-111)."""
    log = cfg.log_path
    print_log(log, "=====================Config=====================")
    for key, value in vars(cfg).items():
        print_log(log, key, ": ", value)
    print_log(log, "======================End=======================")


def show_epoch_info(phase, log_path, info):
    """Per-epoch summary line(s) (reference utils.py:113-129); `info` is the dict train_*/test_* return."""
    print_log(log_path, "")
    head = "%s at epoch #%d" % (phase, info["epoch"])
    print_log(log_path, ("====> " + head) if phase == "Test" else head)
    print_log(log_path, "Group Activity Accuracy: %.2f%%, Loss: %.5f, Using %.1f seconds"
              % (info["activities_acc"], info["loss"], info["time"]))
    if "activities_conf" in info:
        print_log(log_path, info["activities_conf"])
    if "activities_MPCA" in info:
        print_log(log_path, "Activities MPCA:{:.2f}%".format(info["activities_MPCA"]))
    if "MAD" in info:
        print_log(log_path, "MAD:{:.4f}".format(info["MAD"]))
    print_log(log_path, "\n")


_QUIET_CFG_KEYS = frozenset(("num_workers", "use_gpu", "use_multi_gpu", "device_list", "batch_size_test",
                             "test_interval_epoch", "train_random_seed", "result_path", "log_path", "device"))


def log_final_exp_result(log_path, data_path, exp_result):
    """Append the experiment summary to the log and record it in the pickle at `data_path`
    (reference utils.py:131-159)."""
    with open(log_path, "a") as f:
        f.write("\n\n\n")
        print("=====================Config=====================", file=f)
        for key, value in vars(exp_result["cfg"]).items():
            if key not in _QUIET_CFG_KEYS:
                print(key, ": ", value, file=f)
        print("=====================Result======================", file=f)
        print("Best result:", file=f)
        print(exp_result["best_result"], file=f)
        print("Cost total %.4f hours." % (exp_result["total_time"]), file=f)
        print("======================End=======================", file=f)
    with open(data_path, "rb") as f:
        book = pickle.load(f)
    book[exp_result["cfg"].exp_name] = exp_result
    with open(data_path, "wb") as f:
        pickle.dump(book, f)


class AverageMeter(object):
    """Running weighted mean: `.update(val, n)`, `.avg` (reference utils.py:162-179)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class Timer(object):
    """`.timeit()` = seconds since construction or the previous call (reference utils.py:181-191)."""

    def __init__(self):
        self.last_time = time.time()

    def timeit(self):
        now = time.time()
        lap, self.last_time = now - self.last_time, now
        return lap


class ConfusionMeter(object):
    """k x k confusion matrix, rows = ground truth, columns = prediction (reference utils.py:193-276).
    `add` takes class indices [N] or scores / one-hot rows [N,k] (tensors); `value()` returns the int32 counts,
    or row-normalised floats when `normalized`."""

    def __init__(self, k, normalized=False):
        self.k, self.normalized = k, normalized
        self.conf = np.zeros((k, k), dtype=np.int32)

    def reset(self):
        self.conf.fill(0)

    def _indices(self, t, what):
        a = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
        if a.ndim != 1:
            assert a.shape[1] == self.k, "%s does not match the size of the confusion matrix" % what
            if what == "target":
                assert ((a >= 0) & (a <= 1)).all() and (a.sum(1) == 1).all(), "one-hot targets only"
            a = a.argmax(1)
        a = a.astype(np.int64)
        assert a.size == 0 or (a.min() >= 0 and a.max() < self.k), "%s values are not in [0, k)" % what
        return a

    def add(self, predicted, target):
        p, t = self._indices(predicted, "predicted"), self._indices(target, "target")
        assert p.shape[0] == t.shape[0], "number of targets and predicted outputs do not match"
        self.conf += np.bincount(t * self.k + p, minlength=self.k * self.k).reshape(self.k, self.k).astype(np.int32)

    def value(self):
        if not self.normalized:
            return self.conf
        c = self.conf.astype(np.float32)
        return c / c.sum(1).clip(min=1e-12)[:, None]


def MPCA(conf_mat):
    """Mean per-class accuracy in percent from a confusion matrix (reference utils.py:278-289)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        per_class = np.diag(conf_mat).astype(np.float32) / np.sum(conf_mat, axis=1, dtype=np.float32)
    return np.mean(per_class) * 100


def MAC2FLOP(macs, params, module_name=""):
    """Pretty-print a MAC / parameter count as GFLOPs = 2 x MACs (reference utils.py:291-300; no thop needed)."""
    def human(v):
        for unit, div in (("G", 1e9), ("M", 1e6), ("K", 1e3)):
            if v >= div:
                return "%.3f%s" % (v / div, unit)
        return "%.3f" % v
    print("{} MACs: {}  #Params: {}".format(module_name, human(macs), human(params)))
    print("{} GFLOPs: {}G  #Params: {}".format(module_name, 2.0 * macs / 1e9, human(params)))


def _adopt_reference_extras():
    """Pick up public names only the reference's utils.py defines (MADmeter, sincos_encoding_2d, ...) when that
    file is importable from a later sys.path entry; never overrides anything defined above."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    for entry in sys.path:
        cand = os.path.join(entry or ".", "utils.py")
        if os.path.abspath(os.path.dirname(cand)) == here or not os.path.isfile(cand):
            continue
        try:
            with open(cand) as f:
                if "def prep_images" not in f.read():
                    continue
            spec = importlib.util.spec_from_file_location("_reference_utils", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except Exception:
            return
        for name, obj in vars(mod).items():
            if not name.startswith("_") and name not in globals():
                globals()[name] = obj
        return


_adopt_reference_extras()
