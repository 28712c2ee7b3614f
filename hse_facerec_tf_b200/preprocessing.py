"""sklearn.preprocessing.normalize(X, norm='l2') (facerec_test.py:262,265,405) on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib
from .model import _stream_ptr


def normalize(X, norm="l2", device="cuda:0"):
    if norm != "l2":
        raise ValueError("only norm='l2' is on the hot path")
    as_numpy = not isinstance(X, torch.Tensor)
    x = torch.as_tensor(np.asarray(X, dtype=np.float32) if as_numpy else X)
    if not x.is_cuda:
        x = x.to(device)
    x = x.contiguous().float()
    if x.dim() != 2:
        raise ValueError("Expected 2D array")
    y = torch.empty_like(x)
    check(lib.hfr_l2_normalize(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], x.device.index or 0,
                               _stream_ptr(x.device)))
    return y.cpu().numpy() if as_numpy else y
