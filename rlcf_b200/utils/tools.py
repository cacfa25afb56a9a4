"""Seed / meters / top-k accuracy with the reference's names (TPT/utils/tools.py)."""
from __future__ import annotations

import random
from enum import Enum

import numpy as np
import torch


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


class Summary(Enum):
    NONE = 0
    AVERAGE = 1
    SUM = 2
    COUNT = 3


class AverageMeter:
    """Running value / sum / count / average (tools.py:22-59)."""

    def __init__(self, name, fmt=":f", summary_type=Summary.AVERAGE):
        self.name, self.fmt, self.summary_type = name, fmt, summary_type
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(**self.__dict__)

    def summary(self):
        fmt = {Summary.NONE: "", Summary.AVERAGE: "{name} {avg:.3f}", Summary.SUM: "{name} {sum:.3f}",
               Summary.COUNT: "{name} {count:.3f}"}[self.summary_type]
        return fmt.format(**self.__dict__)


class ProgressMeter:
    def __init__(self, num_batches, meters, prefix=""):
        digits = len(str(num_batches // 1))
        self.batch_fmtstr = "[{:" + str(digits) + "d}/" + ("{:" + str(digits) + "d}").format(num_batches) + "]"
        self.meters, self.prefix = meters, prefix

    def display(self, batch):
        print("\t".join([self.prefix + self.batch_fmtstr.format(batch)] + [str(m) for m in self.meters]))

    def display_summary(self):
        print(" ".join([" *"] + [m.summary() for m in self.meters]))


def accuracy(output, target, topk=(1,)):
    """Percentage of rows whose target is among the k highest logits (tools.py:84-98).  With fewer than k classes the
    reference's `topk` raises; here k is clamped to the number of classes (top-5 of a 3-class set is a certain hit)."""
    with torch.no_grad():
        maxk = min(max(topk), output.size(1))
        batch_size = target.size(0)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand(maxk, batch_size))
        return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / batch_size) for k in topk]
