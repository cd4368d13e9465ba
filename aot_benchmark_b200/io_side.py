"""Frame I/O around the hot path on the GPU.

``FramePreprocessor`` replaces ``MultiRestrictSize`` + ``MultiToTensor`` (dataloaders/video_transforms.py:594-715) for the
current image: same constructor arguments, same size rule, same arithmetic (cv2's INTER_CUBIC taps, numpy's normalisation
statements), but the frame is uploaded as uint8 from pinned memory and resized / normalised / transposed by one kernel on the
device, so the DataLoader workers only decode.  ``AsyncMaskWriter`` replaces ``utils.image.save_mask`` (utils/image.py:90-105):
the label map is converted to uint8 on the device, copied into a pinned ring buffer without blocking the stream, and written as
a palette PNG by a small thread pool (the reference converts on the host and starts one thread per frame).
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops


def _cubic_taps(src, dst):
    """cv2 INTER_CUBIC along one axis -> (clamped tap index [dst, 4] int32, weight [dst, 4] float32): sample position
    (d + 0.5) * src / dst - 0.5, Keys kernel with A = -0.75, the fourth weight as 1 - the other three."""
    d = np.arange(dst)
    f = (d + 0.5) * (src / dst) - 0.5
    s = np.floor(f).astype(np.int64)
    x = (f - s).astype(np.float32)
    A = np.float32(-0.75)
    w0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    w1 = ((A + 2) * x - (A + 3)) * x * x + 1
    w2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    w3 = np.float32(1) - w0 - w1 - w2
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, src - 1).astype(np.int32)
    return idx, np.stack([w0, w1, w2, w3], -1).astype(np.float32)


class FramePreprocessor:
    def __init__(self, max_short_edge=None, max_long_edge=800, flip=False, multi_scale=(1.0,), align_corners=True,
                 max_stride=16, device="cuda"):
        self.max_short_edge, self.max_long_edge = max_short_edge, max_long_edge
        self.flip, self.multi_scale = flip, list(multi_scale)
        self.align_corners, self.max_stride = align_corners, max_stride
        self.device = torch.device(device)
        self._taps = {}
        self._pinned = {}

    def target_size(self, h, w, scale):
        """MultiRestrictSize's size rule (video_transforms.py:609-655)."""
        nh, nw = float(h), float(w)
        if self.max_short_edge is not None and min(h, w) > self.max_short_edge:
            r = float(self.max_short_edge) / min(h, w)
            nh, nw = r * nh, r * nw
        if self.max_long_edge is not None and max(nh, nw) > self.max_long_edge:
            r = float(self.max_long_edge) / max(nh, nw)
            nh, nw = r * nh, r * nw
        nh, nw = int(nh * scale), int(nw * scale)
        off = 1 if self.align_corners else 0
        if (nh - off) % self.max_stride != 0:
            nh = int(np.around((nh - off) / self.max_stride) * self.max_stride + off)
        if (nw - off) % self.max_stride != 0:
            nw = int(np.around((nw - off) / self.max_stride) * self.max_stride + off)
        return nh, nw

    def _tables(self, h, w, nh, nw):
        key = (h, w, nh, nw)
        t = self._taps.get(key)
        if t is None:
            iy, cy = _cubic_taps(h, nh)
            ix, cx = _cubic_taps(w, nw)
            t = self._taps[key] = tuple(torch.from_numpy(a).to(self.device) for a in (ix, cx, iy, cy))
        return t

    def upload(self, bgr_u8):
        """uint8 HWC numpy frame (cv2.imread) -> device tensor through a pinned staging buffer (asynchronous copy)."""
        if isinstance(bgr_u8, torch.Tensor):
            return bgr_u8.to(self.device, non_blocking=True)
        pin = self._pinned.get(bgr_u8.shape)
        if pin is None:
            pin = self._pinned[bgr_u8.shape] = torch.empty(bgr_u8.shape, dtype=torch.uint8).pin_memory()
        pin.numpy()[...] = bgr_u8
        return pin.to(self.device, non_blocking=True)

    def __call__(self, bgr_u8):
        """-> list of float32 tensors [1, 3, h, w] on the device, one per (scale, flip) in MultiRestrictSize's order."""
        img = self.upload(bgr_u8)
        h, w = int(img.shape[0]), int(img.shape[1])
        outs = []
        for scale in self.multi_scale:
            nh, nw = self.target_size(h, w, scale)
            taps = None if (nh, nw) == (h, w) else self._tables(h, w, nh, nw)
            for fl in ((False, True) if self.flip else (False,)):
                out = torch.empty((1, 3, nh, nw), dtype=torch.float32, device=self.device)
                ops.preprocess_bgr_u8(img, out, taps, fl)
                outs.append(out)
        return outs


def davis_palette():
    """utils/image.py:6-59 from its rule: 22 bit-interleaved VOC colours (with 191 where VOC has 192), then greys."""
    pal = []
    for i in range(22):
        c, rgb = i, [0, 0, 0]
        for j in range(8):
            for ch in range(3):
                rgb[ch] |= ((c >> ch) & 1) << (7 - j)
            c >>= 3
        pal += [191 if v == 192 else v for v in rgb]
    for i in range(22, 256):
        pal += [i, i, i]
    return pal


class AsyncMaskWriter:
    """save_mask(mask_tensor, path, squeeze_idx) without stalling the propagation loop: uint8 conversion on the device, copy
    into a ring of pinned buffers on the caller's stream, PNG encoding on worker threads."""

    def __init__(self, workers=4, ring=16):
        self._pool = ThreadPoolExecutor(max_workers=workers)
        self._ring = [None] * ring
        self._busy = [threading.Event() for _ in range(ring)]
        for e in self._busy:
            e.set()
        self._next = 0
        self._palette = davis_palette()
        self._futures = []

    def save(self, mask_tensor, path, squeeze_idx=None):
        m = mask_tensor
        slot = self._next
        self._next = (self._next + 1) % len(self._ring)
        self._busy[slot].wait()                            # the worker that used this pinned buffer has finished
        self._busy[slot].clear()
        shape = tuple(int(s) for s in m.shape[-2:])
        buf = self._ring[slot]
        if buf is None or tuple(buf.shape) != shape:
            buf = torch.empty(shape, dtype=torch.uint8)
            buf = self._ring[slot] = buf.pin_memory() if m.is_cuda else buf
        if m.is_cuda:
            u8 = torch.empty(shape, dtype=torch.uint8, device=m.device)
            ops.label_to_u8(m.reshape(shape).float().contiguous(), u8)
            buf.copy_(u8, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        else:
            buf.copy_(m.reshape(shape).to(torch.uint8))
            done = None
        self._futures.append(self._pool.submit(self._write, slot, done, path, squeeze_idx))

    def _write(self, slot, done, path, squeeze_idx):
        from PIL import Image
        try:
            if done is not None:
                done.synchronize()
            mask = self._ring[slot].numpy().copy()
        finally:
            self._busy[slot].set()
        if squeeze_idx is not None:                        # utils/image.py:91-97: compact ids -> original object ids
            lut = np.zeros(256, dtype=np.uint8)
            for idx in range(1, len(squeeze_idx)):
                lut[idx] = squeeze_idx[idx]
            mask = lut[mask]
        im = Image.fromarray(mask).convert('P')
        im.putpalette(self._palette)
        im.save(path)

    def flush(self):
        for f in self._futures:
            f.result()
        self._futures = []

    def close(self):
        self.flush()
        self._pool.shutdown()
