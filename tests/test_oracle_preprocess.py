"""The preprocessing oracle (oracle/preprocess_oracle.py) pinned against the installed OpenCV itself (cv2 is the
reference's own dependency, requirements.txt:3) and against a committed fixture made from it.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import preprocess_oracle as P

cv2 = pytest.importorskip('cv2')

CASES = [(120, 173, 0.33), (97, 64, 0.25), (300, 211, 0.33), (64, 64, 0.5), (333, 517, 0.2), (50, 77, 0.25), (51, 79, 0.33),
         (408, 533, 0.33)]


@pytest.mark.parametrize('H,W,f', CASES)
def test_resize_pad_against_live_cv2(H, W, f):
    rng = np.random.RandomState(H * 7 + W)
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    ref = cv2.resize(img, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_AREA)
    got = P.resize_area(img, f)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    mask = rng.randint(0, 6, (H, W)).astype(np.uint8)
    assert np.array_equal(P.resize_nearest(mask, f), cv2.resize(mask, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_NEAREST))
    refp = cv2.copyMakeBorder(ref, 0, (-ref.shape[0]) % 32, 0, (-ref.shape[1]) % 32, cv2.BORDER_CONSTANT)
    assert np.array_equal(P.pad_bottom_right(got, 32), refp)


def test_preprocess_fixture():
    """tests/golden/preprocess.npz: input + cv2 outputs recorded by oracle/gen_golden.py::gen_preprocess."""
    g = load_golden('preprocess')
    for tag in ('a', 'b'):
        img, f = g[f'img_{tag}'], float(g[f'factor_{tag}'])
        assert np.array_equal(P.resize_area(img, f), g[f'resized_{tag}'])
        out = P.preprocess_scene(img, f, 32)
        assert out.dtype == np.float32 and np.array_equal(out, g[f'chw_{tag}'])
    assert np.array_equal(P.preprocess_scene(g['mask'], 0.33, 32, seg_mask=True), g['onehot'])
