"""SURVEY 8f rank 2 on the GPU: the fused scene-image preprocessing kernel (csrc/preprocess.cu) through the C ABI against
the oracle (itself pinned against the installed cv2, tests/test_oracle_preprocess.py) and the cv2-recorded fixture.
Byte / integer work: BIT-EXACT, float32 normalised output included (same float64 arithmetic)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import preprocess_oracle as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('H,W,f', [(120, 173, 0.33), (97, 64, 0.25), (64, 64, 0.5), (333, 517, 0.2), (51, 79, 0.33),
                                   (1088, 1424, 0.33), (1088, 1424, 0.25)])
def test_fused_preprocess_bit_exact(cuda_device, H, W, f):
    from motion_style_transfer_b200.utils import image_utils as U
    rng = np.random.RandomState(H + W)
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    got = U.preprocess_scene_image(img, f, 32).cpu().numpy()
    ref = P.preprocess_scene(img, f, 32)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got, ref)
    mask = rng.randint(0, 6, (H, W)).astype(np.uint8)
    assert np.array_equal(U.preprocess_scene_image(mask, f, 32, seg_mask=True).cpu().numpy(),
                          P.preprocess_scene(mask, f, 32, seg_mask=True))


def test_reference_named_helpers_and_fixture(cuda_device):
    """resize / pad / preprocess_image_for_segmentation keep the reference's in-place dict semantics
    (image_utils.py:66-107) and reproduce the cv2-recorded fixture."""
    from motion_style_transfer_b200.utils import image_utils as U
    g = load_golden('preprocess')
    images = {'a': g['img_a'].copy(), 'b': g['img_b'].copy()}
    factors = {'a': float(g['factor_a']), 'b': float(g['factor_b'])}
    for k in ('a', 'b'):
        one = {k: images[k]}
        U.resize(one, factors[k])
        assert np.array_equal(one[k], g[f'resized_{k}'])
        U.pad(one, 32)
        assert one[k].shape[0] % 32 == 0 and one[k].shape[1] % 32 == 0
        U.preprocess_image_for_segmentation(one)
        assert np.array_equal(one[k].cpu().numpy(), g[f'chw_{k}'])
    m = {'m': g['mask'].copy()}
    U.resize(m, 0.33, seg_mask=True)
    U.pad(m, 32)
    U.preprocess_image_for_segmentation(m, seg_mask=True)
    assert np.array_equal(m['m'].cpu().numpy(), g['onehot'])
    with pytest.raises(NotImplementedError):
        U.preprocess_image_for_segmentation({'x': g['img_a']}, encoder='resnet34')


def test_semantic_map_is_cached_per_scene(cuda_device):
    """YNet.segmentation_cached: the backbone runs once per (scene image, backbone weights)."""
    from motion_style_transfer_b200.models.ynet import YNet
    m = YNet(obs_len=5, pred_len=6, segmentation_model_fp=None, encoder_channels=[8, 8, 16, 16, 16],
             decoder_channels=[16, 16, 16, 8, 8], n_waypoints=2, train_net='mosa_1', position=[0, 1, 2, 3, 4],
             network='original').cuda()
    calls = []

    class Seg(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def forward(self, x):
            calls.append(1)
            return x * self.w

    m.semantic_segmentation = Seg().cuda()
    img = torch.rand(1, 6, 32, 32, device='cuda')
    a = m.segmentation_cached('s0', img)
    b = m.segmentation_cached('s0', img)
    assert a is b and len(calls) == 1
    m.segmentation_cached('s1', img.clone())
    assert len(calls) == 2
    img.mul_(0.5)                                   # the scene image changed in place
    m.segmentation_cached('s0', img)
    with torch.no_grad():
        m.semantic_segmentation.w.add_(1.0)         # the backbone changed
    m.segmentation_cached('s0', img)
    assert len(calls) == 4
