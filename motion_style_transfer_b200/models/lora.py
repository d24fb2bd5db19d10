"""LoRA-adapted 3x3 convolution with the parameter layout of ``loralib==0.1.1``.

The reference builds its MoSA layers with ``lora.Conv2d(in, out, kernel_size=3, r=rank, stride=1,
padding=1)`` (models/ynet.py:143-144); checkpoints therefore carry the flat keys
``<conv>.{weight,bias,lora_A,lora_B}`` (evaluator/analyze_lora_importance.py:74-79).  This module
is a parameter container with exactly those names, shapes and initialisers; its arithmetic
(W + (B @ A).view(W.shape) / r, never merged on this path) is executed by the CUDA engine
(``ynet_lora_fold`` + conv kernels), not by torch.
"""
import math

import torch
import torch.nn as nn


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, r=0, lora_alpha=1, lora_dropout=0.0,
                 merge_weights=True, **kwargs):
        if not isinstance(kernel_size, int):
            raise TypeError('lora Conv2d expects an int kernel_size')
        self._lora_built = False
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        self.r = r
        self.lora_alpha = lora_alpha
        self.merged = False
        self.merge_weights = merge_weights
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r * kernel_size, in_channels * kernel_size)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_channels * kernel_size, r * kernel_size)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
        self._lora_built = True
        self.reset_parameters()

    def reset_parameters(self):
        super().reset_parameters()
        if getattr(self, '_lora_built', False) and hasattr(self, 'lora_A'):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def forward(self, x):  # pragma: no cover - the engine runs the arithmetic
        raise RuntimeError('motion_style_transfer_b200: convolutions run through the CUDA engine '
                           '(YNet.pred_features / pred_goal / pred_traj), not through nn.Module.forward')
