"""TEST INFRASTRUCTURE ONLY: readers for the compact full-geometry fixtures (tests/golden/full_*.pt) written by
oracle/gen_golden.py (which imports the real reference and therefore cannot be imported on the GPU box)."""
import zlib

import numpy as np
import torch


def load_full_labels(g):
    """-> list of [1,1,oh,ow] float label maps (the real reference's argmax masks, one per propagated frame)."""
    arr = np.frombuffer(zlib.decompress(g["ref_labels_zlib"]), dtype=np.uint8).reshape(g["ref_labels_shape"])
    return [torch.from_numpy(arr[i].copy()).float()[None, None] for i in range(arr.shape[0])]
