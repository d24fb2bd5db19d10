"""Restatement of ``loralib==0.1.1``'s ``Conv2d`` (test infrastructure only).

The reference depends on the un-vendored package ``loralib==0.1.1``
(/root/reference/requirements.txt:11) and uses exactly one symbol from it:
``lora.Conv2d(in, out, kernel_size=3, r=rank, stride=1, padding=1)``
(/root/reference/models/ynet.py:143-144).  The package is not installed in this
image and cannot be fetched, so its published algorithm is restated here:

* parameters ``lora_A (r*k, C_in*k)`` (kaiming-uniform, a=sqrt(5)) and
  ``lora_B (C_out*k, r*k)`` (zeros) live directly on the conv module
  (flat 0.1.1 key layout, confirmed by
  /root/reference/evaluator/analyze_lora_importance.py:74-79);
* ``scaling = lora_alpha / r`` with ``lora_alpha = 1``;
* ``weight.requires_grad = False``;
* forward, when not merged:
  ``conv2d(x, weight + (lora_B @ lora_A).view(weight.shape) * scaling, bias)``;
* ``eval()`` merges, ``train(mode)`` un-merges -- but a parent module's
  ``.eval()`` reaches children as ``.train(False)``, so on the reference's path
  (``model.eval()`` in utils/evaluate.py:68) weights are never merged.

PARITY UNPINNED: this file cannot be diffed against the real wheel offline.
"""
import math
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, r=0, lora_alpha=1,
                 lora_dropout=0.0, merge_weights=True, **kwargs):
        assert isinstance(kernel_size, int)
        self._lora_ready = False
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, **kwargs)
        self.r = r
        self.lora_alpha = lora_alpha
        self.merged = False
        self.merge_weights = merge_weights
        if r > 0:
            self.lora_A = nn.Parameter(
                self.weight.new_zeros((r * kernel_size, in_channels * kernel_size)))
            self.lora_B = nn.Parameter(
                self.weight.new_zeros((out_channels * kernel_size, r * kernel_size)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
        self._lora_ready = True
        self.reset_parameters()

    def reset_parameters(self):
        nn.Conv2d.reset_parameters(self)
        if getattr(self, '_lora_ready', False) and hasattr(self, 'lora_A'):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def _delta(self):
        return (self.lora_B @ self.lora_A).view(self.weight.shape) * self.scaling

    def train(self, mode=True):
        nn.Conv2d.train(self, mode)
        if self.merge_weights and self.merged:
            self.weight.data -= self._delta()
            self.merged = False
        return self

    def eval(self):
        nn.Conv2d.eval(self)
        if self.merge_weights and not self.merged:
            self.weight.data += self._delta()
            self.merged = True
        return self

    def forward(self, x):
        if self.r > 0 and not self.merged:
            return F.conv2d(x, self.weight + self._delta(), self.bias,
                            self.stride, self.padding, self.dilation, self.groups)
        return nn.Conv2d.forward(self, x)


def install_as_loralib():
    """Register this restatement as the importable module ``loralib``."""
    if 'loralib' in sys.modules:
        return sys.modules['loralib']
    mod = types.ModuleType('loralib')
    mod.Conv2d = Conv2d
    mod.__doc__ = 'restatement of loralib==0.1.1 (oracle/loralib_restatement.py)'
    sys.modules['loralib'] = mod
    return mod
