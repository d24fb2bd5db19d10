"""SURVEY 8f rank 4, augmentation (utils/data_utils.py:115-233): the drop-in ``augment_data`` against a fixture recorded
from the LIVE reference (oracle/gen_golden.py::gen_augment) -- trajectories bit-identical on the host; the eight image
views of a scene, read by the preprocessing kernel from ONE stored image, bit-identical on the GPU.
"""
import numpy as np
import pandas as pd
import pytest
import torch

from conftest import load_golden


def _frame(g):
    return pd.DataFrame({'frame': np.tile(np.arange(4), len(g['in/x']) // 4), 'trackId': g['in/metaId'], 'x': g['in/x'],
                         'y': g['in/y'], 'sceneId': g['in/sceneId'].astype(object), 'metaId': g['in/metaId']})


def _images(g):
    return {k[4:]: g[k] for k in g.files if k.startswith('img/')}


def test_augment_data_trajectories_bit_identical():
    from motion_style_transfer_b200.utils.image_utils import augment_data
    g = load_golden('augment')
    data, views = augment_data(_frame(g), _images(g))
    assert list(data.sceneId.values) == list(g['out/sceneId'])
    assert np.array_equal(data.metaId.values, g['out/metaId'])
    assert np.array_equal(data.x.values, g['out/x']) and np.array_equal(data.y.values, g['out/y'])      # float64, bit for bit
    assert len(data) == 8 * len(g['in/x'])
    assert set(views) == {k[4:] for k in g.files if k.startswith('chw/')}
    assert views['sA_rot270_fliplr'] == ('sA', 7) and views['s_B'] == ('s_B', 0) and views['s_B_fliplr'] == ('s_B', 4)


@pytest.mark.gpu
def test_augmented_views_from_one_stored_image_bit_exact(cuda_device):
    """rot90 x k + fliplr + resize (INTER_AREA) + pad + normalise of the reference == one oriented preprocessing launch."""
    from motion_style_transfer_b200.utils.image_utils import augment_data, preprocess_scene_image
    g = load_golden('augment')
    images = _images(g)
    _, views = augment_data(_frame(g), images)
    for view, (base, orient) in views.items():
        got = preprocess_scene_image(images[base], float(g['factor']), 32, orient=orient).cpu().numpy()
        ref = g['chw/' + view]
        assert got.shape == ref.shape, view
        assert np.array_equal(got, ref), view
    # the non-integer scale path (area tables) in a rotated frame: against OpenCV itself
    import cv2
    img = np.random.RandomState(3).randint(0, 256, (97, 131, 3)).astype(np.uint8)
    from oracle import preprocess_oracle as P
    for orient in range(8):
        v = img
        for _ in range(orient & 3):
            v = cv2.rotate(v, cv2.ROTATE_90_COUNTERCLOCKWISE)
        if orient >> 2:
            v = cv2.flip(v, 1)
        r = cv2.resize(v, (0, 0), fx=0.33, fy=0.33, interpolation=cv2.INTER_AREA)
        p = cv2.copyMakeBorder(r, 0, (-r.shape[0]) % 32, 0, (-r.shape[1]) % 32, cv2.BORDER_CONSTANT)
        ref = (((p / 255.0) - P.IMAGENET_MEAN) / P.IMAGENET_STD).transpose(2, 0, 1).astype('float32')
        got = preprocess_scene_image(img, 0.33, 32, orient=orient).cpu().numpy()
        assert np.array_equal(got, ref), orient


@pytest.mark.gpu
def test_prepare_data_with_augmentation(cuda_device, tmp_path):
    """YNetTrainer.prepare_data(augment=True) (trainer.py:566-571): 8x scenes, device tensors, loader over all of them."""
    import cv2
    from motion_style_transfer_b200.models.trainer import YNetTrainer
    g = load_golden('augment')
    for scene, im in _images(g).items():
        (tmp_path / scene).mkdir()
        cv2.imwrite(str(tmp_path / scene / 'reference.png'), im)
    params = dict(obs_len=2, pred_len=2, segmentation_model_fp=None, use_features_only=False, n_semantic_classes=6,
                  encoder_channels=[8, 8, 16, 16, 16], decoder_channels=[16, 16, 16, 8, 8], waypoints=[1],
                  train_net='mosa_1', position=[0, 1, 2, 3, 4], network='original', n_fusion=None, resize_factor=0.5)
    t = YNetTrainer(params, device=torch.device('cuda'))
    images, loader, homo = t.prepare_data(_frame(g), str(tmp_path), 'ind-dataset-v1.0', 'val', 2, 2, 0.5, False, augment=True)
    assert homo is None and len(images) == 16 and len(loader) == 16
    for view in images:
        assert np.array_equal(images[view].cpu().numpy(), g['chw/' + view]), view
    seen = [scene for _, _, scene in loader]
    assert sorted(seen) == sorted(images)
