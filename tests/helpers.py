"""Shared helpers for the parity tests (CPU side: oracle; GPU side: product through the C ABI)."""
import ast

import numpy as np
import torch

from conftest import load_golden, golden_state_dict

SMALL_ENC = [8, 8, 16, 16, 16]
SMALL_DEC = [16, 16, 16, 8, 8]


def build_product_model(sd, obs, pred, n_wp, network='original', position=(0, 1, 2, 3, 4), n_fusion=None,
                        train_net='mosa_1', enc=SMALL_ENC, dec=SMALL_DEC, device='cuda'):
    from motion_style_transfer_b200.models.ynet import YNet
    m = YNet(obs_len=obs, pred_len=pred, segmentation_model_fp=None, encoder_channels=list(enc),
             decoder_channels=list(dec), n_waypoints=n_wp, train_net=train_net, position=list(position),
             network=network, n_fusion=n_fusion)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


class ReplayRng:
    """Feeds randoms recorded in a fixture (the ones the live reference consumed)."""

    def __init__(self, g):
        self.g = g

    def uniforms(self, rows, n, device):
        return torch.from_numpy(self.g['uniforms']).to(device)

    def exponentials(self, rows, S, device):
        return torch.from_numpy(self.g['expo']).to(device)


def eval_cfg(g):
    return ast.literal_eval(str(g['cfg']))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


ADAPTER_TAGS = ['parallelLayer_3x3', 'parallelLayer_1x1_3x3', 'serialLayer', 'serial_block', 'parallel_block_1x1',
                'parallel_block_1x1_3x3']


def build_adapter_model(g, device='cpu'):
    """Drop-in YNet of an ``adapter_<tag>`` fixture (oracle/gen_golden.py::gen_adapters): decoders = the default
    initialisation under the fixture's seed (bit-identical to the reference's, same modules in the same order), encoder =
    the fixture's tensors."""
    from motion_style_transfer_b200.models.ynet import YNet
    torch.manual_seed(int(g['seed']))
    m = YNet(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=list(SMALL_ENC),
             decoder_channels=list(SMALL_DEC), n_waypoints=2, train_net=str(g['train_net']),
             position=[int(p) for p in g['position']], network='original')
    sd = golden_state_dict(g)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith('encoder.') for k in missing)
    return m.to(device).eval()


class TinySeg(torch.nn.Module):
    """Stand-in for the pickled smp U-Net the reference loads from ``<data_dir>/<dataset>/<dataset>_segmentation.pth``
    (ynet.py:495-507): a whole pickled module, 3 -> n_classes channels at the input resolution."""

    def __init__(self, n_classes=6):
        super().__init__()
        self.head = torch.nn.Conv2d(3, n_classes, 3, padding=1)

    def forward(self, x):
        return self.head(x)
