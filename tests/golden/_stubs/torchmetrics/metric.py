"""Minimal stand-in for torchmetrics.metric.Metric: state registration, forward = update + compute, reset."""
import torch


class Metric(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = default.clone()
        setattr(self, name, default.clone())

    def forward(self, *a, **k):
        self.update(*a, **k)
        return self.compute()

    def reset(self):
        for n, d in self._defaults.items():
            setattr(self, n, d.clone())
