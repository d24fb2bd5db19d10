"""CPU restatement of the scene-image preprocessing in front of the hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

  resize   utils/image_utils.py:85-92    cv2.resize(fx = fy = factor, INTER_AREA | INTER_NEAREST for segmentation masks)
  pad      utils/image_utils.py:95-107   cv2.copyMakeBorder(bottom / right, BORDER_CONSTANT 0) up to a multiple of 32
  preprocess_image_for_segmentation  utils/image_utils.py:66-82   smp preprocessing (x / 255 - mean) / std, HWC -> CHW
                                                                  float32; one-hot for segmentation masks

cv2.resize lives in OpenCV (opencv-python, requirements.txt:3; 4.13 in this image), not under /root/reference: its
INTER_AREA algorithm (imgproc/src/resize.cpp: computeResizeAreaTab, ResizeArea_Invoker, ResizeAreaFast_Invoker) is restated
here and PINNED against the installed cv2 itself (tests/test_oracle_preprocess.py, live, plus tests/golden/preprocess.npz).
segmentation_models_pytorch==0.1.0 (requirements.txt:7) is absent; its preprocessing for encoder 'resnet101' / weights
'imagenet' is the published constant set below (input_space RGB, input_range [0, 1]) -- parity unpinned for that constant
set, pinned for the arithmetic (float64, then astype float32).
"""
import numpy as np

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])
IMAGENET_STD = np.array([0.229, 0.224, 0.225])


def area_tab(ssize, dsize, scale):
    """cv::computeResizeAreaTab: [(dst index, src index, float32 weight)] in accumulation order."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = int(np.ceil(fsx1)), int(np.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_plan(H, W, factor):
    """Output size and the path cv::resize(INTER_AREA, fx = fy = factor, dsize = (0, 0)) takes: dsize = cvRound(size * f),
    scale = 1 / f (NOT size ratio); integer scales with an exact fit take the 'fast' integer path."""
    dh, dw = int(np.rint(H * factor)), int(np.rint(W * factor))
    scale = 1.0 / factor
    isc = int(np.rint(scale))
    fast = abs(scale - isc) < np.finfo(np.float64).eps and dh * isc <= H and dw * isc <= W
    return dh, dw, scale, (isc if fast else 0)


def resize_area(img, factor):
    """image_utils.py:91-92: cv2.resize(image, (0, 0), fx=factor, fy=factor, interpolation=cv2.INTER_AREA), uint8 HWC."""
    img = np.asarray(img)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    H, W, C = img.shape
    dh, dw, scale, isc = resize_plan(H, W, factor)
    if isc:
        acc = np.zeros((dh, dw, C), np.int64)
        src = img.astype(np.int64)
        for ky in range(isc):
            for kx in range(isc):
                acc += src[ky:dh * isc:isc, kx:dw * isc:isc]
        if isc == 2:
            out = ((acc + 2) >> 2).astype(np.uint8)                   # ResizeAreaFastVec: round half up
        else:
            out = np.clip(np.rint(acc.astype(np.float32) * np.float32(1.0 / (isc * isc))), 0, 255).astype(np.uint8)
    else:
        xt, yt = area_tab(W, dw, scale), area_tab(H, dh, scale)
        src = img.astype(np.float32)
        bx = np.zeros((H, dw, C), np.float32)
        for dx, s, a in xt:                                           # x pass: buf[dx] += S[sx] * alpha   (float32, no FMA)
            bx[:, dx] = bx[:, dx] + src[:, s] * a
        acc = np.zeros((dh, dw, C), np.float32)
        for dy, s, b in yt:                                           # y pass: sum[dx] += buf[dx] * beta
            acc[dy] = acc[dy] + bx[s] * b
        out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)          # saturate_cast<uchar>: cvRound, half to even
    return out[:, :, 0] if squeeze else out


def resize_nearest(img, factor):
    """image_utils.py:88-90 (segmentation masks): cv2.INTER_NEAREST: src index = min(floor(dst * (1 / f)), size - 1)."""
    img = np.asarray(img)
    H, W = img.shape[:2]
    dh, dw = int(np.rint(H * factor)), int(np.rint(W * factor))
    ys = np.minimum(np.floor(np.arange(dh) * (1.0 / factor)).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(dw) * (1.0 / factor)).astype(np.int64), W - 1)
    return img[ys][:, xs]


def pad_bottom_right(img, division_factor=32):
    """image_utils.py:95-107."""
    H, W = img.shape[:2]
    Hn = int(np.ceil(H / division_factor) * division_factor)
    Wn = int(np.ceil(W / division_factor) * division_factor)
    widths = ((0, Hn - H), (0, Wn - W)) + ((0, 0),) * (img.ndim - 2)
    return np.pad(img, widths, mode='constant')


def normalise_imagenet(img_u8):
    """image_utils.py:79-82 with smp's preprocess_input for resnet101 / imagenet: float64 arithmetic, CHW float32."""
    x = img_u8 / 255.0
    x = (x - IMAGENET_MEAN) / IMAGENET_STD
    return x.transpose(2, 0, 1).astype('float32')


def one_hot(mask, classes=6):
    """image_utils.py:76-78."""
    im = np.stack([(mask == v) for v in range(classes)], axis=-1)
    return im.transpose(2, 0, 1).astype('float32')


def preprocess_scene(img_u8, factor, division_factor=32, seg_mask=False, classes=6):
    """trainer.py:578-582: resize -> pad -> preprocess_image_for_segmentation for ONE scene image."""
    if seg_mask:
        return one_hot(pad_bottom_right(resize_nearest(img_u8, factor), division_factor), classes)
    return normalise_imagenet(pad_bottom_right(resize_area(img_u8, factor), division_factor))
