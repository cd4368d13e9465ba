"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's frame I/O around the hot path.

* restrict_size      -- MultiRestrictSize's size rule (dataloaders/video_transforms.py:594-655)
* resize_cubic       -- cv2.resize(float32 image, INTER_CUBIC) as the reference calls it (:667-670): 4-tap Keys kernel with
                        A = -0.75, sample position (d + 0.5) * scale - 0.5, replicated borders, horizontal then vertical pass
                        in float32.  cv2 is a third-party dependency of the reference (opencv-python, no pinned version); this
                        restatement is pinned against cv2 4.13 in tests/test_cpu_io_side.py (max |d| 4e-4 on a 0..255 scale)
* to_tensor          -- MultiToTensor (:693-715): / 255. in float32, - mean and / std through float64, HWC -> CHW
* davis_palette      -- utils/image.py:6-59 rebuilt from its rule (VOC bit-interleaved colours with 191 for 192, then greys)
* save_mask          -- utils/image.py:90-100 (_save_mask): optional id remap, PIL 'P' image + palette
"""
import numpy as np
import torch


def restrict_size(h, w, max_short_edge=None, max_long_edge=800, scale=1.0, align_corners=True, max_stride=16):
    sc = 1.
    if max_short_edge is not None:
        short = w if h > w else h
        if short > max_short_edge:
            sc *= float(max_short_edge) / short
    new_h, new_w = sc * h, sc * w
    sc = 1.
    if max_long_edge is not None:
        long_edge = new_h if new_h > new_w else new_w
        if long_edge > max_long_edge:
            sc *= float(max_long_edge) / long_edge
    new_h, new_w = sc * new_h, sc * new_w
    new_h, new_w = int(new_h * scale), int(new_w * scale)
    if align_corners:
        if (new_h - 1) % max_stride != 0:
            new_h = int(np.around((new_h - 1) / max_stride) * max_stride + 1)
        if (new_w - 1) % max_stride != 0:
            new_w = int(np.around((new_w - 1) / max_stride) * max_stride + 1)
    else:
        if new_h % max_stride != 0:
            new_h = int(np.around(new_h / max_stride) * max_stride)
        if new_w % max_stride != 0:
            new_w = int(np.around(new_w / max_stride) * max_stride)
    return new_h, new_w


def cubic_taps(src, dst):
    """-> (index [dst, 4] int32 clamped to the image, weight [dst, 4] float32) of cv2's INTER_CUBIC along one axis."""
    scale = src / dst
    d = np.arange(dst)
    f = (d + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    x = (f - s).astype(np.float32)
    A = np.float32(-0.75)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c3 = np.float32(1) - c0 - c1 - c2
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, src - 1).astype(np.int32)
    return idx, np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def resize_cubic(img, Ho, Wo):
    """img float32 [H, W, C] -> [Ho, Wo, C]."""
    H, W, C = img.shape
    iy, cy = cubic_taps(H, Ho)
    ix, cx = cubic_taps(W, Wo)
    tmp = np.zeros((H, Wo, C), np.float32)
    for k in range(4):
        tmp += img[:, ix[:, k], :] * cx[None, :, k, None]
    out = np.zeros((Ho, Wo, C), np.float32)
    for k in range(4):
        out += tmp[iy[:, k]] * cy[:, k, None, None]
    return out


MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def to_tensor(img):
    """float32 HWC (channel order as loaded: the reference keeps cv2's BGR) -> float32 tensor [3, H, W]."""
    tmp = img / 255.
    tmp -= MEAN
    tmp /= STD
    return torch.from_numpy(np.ascontiguousarray(tmp.transpose((2, 0, 1))))


def preprocess(bgr_u8, max_short_edge=None, max_long_edge=800, scale=1.0, align_corners=True, max_stride=16, flip=False):
    """cv2.imread output (uint8 HWC) -> the tensor MultiRestrictSize + MultiToTensor hand to the engine, [3, h, w]."""
    img = np.array(bgr_u8, dtype=np.float32)                      # eval_datasets.py:60-61
    h, w = img.shape[:2]
    nh, nw = restrict_size(h, w, max_short_edge, max_long_edge, scale, align_corners, max_stride)
    if (nh, nw) != (h, w):
        img = resize_cubic(img, nh, nw)
    if flip:
        img = img[:, ::-1].copy()
    return to_tensor(img)


def davis_palette():
    pal = []
    for i in range(22):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        pal += [191 if v == 192 else v for v in (r, g, b)]
    for i in range(22, 256):
        pal += [i, i, i]
    return pal


def save_mask(mask_u8, path, squeeze_idx=None):
    from PIL import Image
    mask = mask_u8
    if squeeze_idx is not None:
        out = mask * 0
        for idx in range(1, len(squeeze_idx)):
            out += ((mask == idx) * squeeze_idx[idx]).astype(np.uint8)
        mask = out
    im = Image.fromarray(mask).convert('P')
    im.putpalette(davis_palette())
    im.save(path)
