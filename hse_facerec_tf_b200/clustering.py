"""Distance matrices for the reference's clustering scripts (scope row 8f-3), on the GPU; the linkage stays on the host.

  pairwise_distances(X[, Y])         sklearn.metrics.pairwise_distances(X_norm)     facial_clustering_test.py:396-400
  album_distance_matrix(f, y, b)     perform_clustering's feature_distance matrix   process_photos.py:46-56
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib
from .model import _stream_ptr


def _dev_f32(a, device):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32))
    return t.to(device, torch.float32).contiguous()


def pairwise_distances(X, Y=None, device="cuda:0"):
    """Euclidean distance matrix [n, m] (float32); Y=None: Y = X with exact zeros on the diagonal, as sklearn returns.
    numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    as_numpy = not isinstance(X, torch.Tensor)
    x = _dev_f32(X, device)
    y = None if Y is None else _dev_f32(Y, x.device)
    if x.dim() != 2 or (y is not None and (y.dim() != 2 or y.shape[1] != x.shape[1])):
        raise ValueError("X and Y must be 2-D with the same number of features")
    n, m = x.shape[0], (x.shape[0] if y is None else y.shape[0])
    out = torch.empty((n, m), dtype=torch.float32, device=x.device)
    if n and m:
        with torch.cuda.device(x.device):
            check(lib.hfr_pairwise_dist(x.data_ptr(), n, None if y is None else y.data_ptr(), m, x.shape[1], None, None, None,
                                        None, 0.0, out.data_ptr(), x.device.index or 0, _stream_ptr(x.device)))
    return out.cpu().numpy() if as_numpy else out


def album_distance_matrix(features, photo_years, born_years, age_weight=0.1, device="cuda:0"):
    """dist_matrix of process_photos.py:54-55: ||f_i - f_j|| + age_weight * (a_i - a_j)^2 / (a_i + a_j), clipped at 0,
    where a = max(year_i, year_j) - born_year (the apparent ages of both faces in the later photo's year).
    features [n, D]; photo_years / born_years [n].  Returns a numpy float32 [n, n]."""
    x = _dev_f32(features, device)
    yr, bn = _dev_f32(photo_years, x.device), _dev_f32(born_years, x.device)
    n = x.shape[0]
    if x.dim() != 2 or yr.shape != (n,) or bn.shape != (n,):
        raise ValueError("features must be [n, D], photo_years and born_years [n]")
    out = torch.empty((n, n), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            check(lib.hfr_pairwise_dist(x.data_ptr(), n, None, n, x.shape[1], yr.data_ptr(), bn.data_ptr(), None, None,
                                        float(age_weight), out.data_ptr(), x.device.index or 0, _stream_ptr(x.device)))
    return out.cpu().numpy()
