"""PCA step of the reference's '1-NN+PCA' / '3-NN+PCA' pipelines (facerec_test.py:271-274,417-422; scope row 8f-2).

fit is scikit-learn's own (the SVD of a training matrix is a one-off on the host, exactly the reference's); transform -
the per-query work - runs on the GPU:  X @ components_.T - mean_ @ components_.T  in fp32 (the library's SIMT GEMM, no
tensor-core rounding), optionally whitened, as sklearn's PCA._transform does.  A drop-in for sklearn.decomposition.PCA
inside a Pipeline; with output="torch" the projected batch stays on the device for the KNeighborsClassifier step.
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.decomposition import PCA as _SkPCA

from ._lib import check, lib
from .model import _stream_ptr


def _project(x: torch.Tensor, comp: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """y[n, c] = x[n, d] @ comp[c, d].T + bias[c] on x's CUDA device, fp32 (hfr_op_gemm_bias_act, precision 0)."""
    n, d = x.shape
    c = comp.shape[0]
    y = torch.empty((n, c), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            check(lib.hfr_op_gemm_bias_act(x.data_ptr(), comp.data_ptr(), bias.data_ptr(), None, y.data_ptr(), n, c, d, 0, 0,
                                           x.device.index or 0, _stream_ptr(x.device)))
    return y


class PCA(_SkPCA):
    def __init__(self, n_components=None, *, copy=True, whiten=False, svd_solver="auto", tol=0.0, iterated_power="auto",
                 n_oversamples=10, power_iteration_normalizer="auto", random_state=None, device="cuda:0", output="numpy"):
        super().__init__(n_components, copy=copy, whiten=whiten, svd_solver=svd_solver, tol=tol,
                         iterated_power=iterated_power, n_oversamples=n_oversamples,
                         power_iteration_normalizer=power_iteration_normalizer, random_state=random_state)
        self.device = device
        self.output = output

    def fit(self, X, y=None):
        X = X.cpu().numpy() if isinstance(X, torch.Tensor) else X
        super().fit(X, y)
        self._dev_cache = None
        return self

    def fit_transform(self, X, y=None):
        return self.fit(X, y).transform(X)

    def _device_params(self):
        cache = getattr(self, "_dev_cache", None)
        if cache is None:
            comp = np.ascontiguousarray(self.components_, dtype=np.float32)
            mean = np.asarray(self.mean_, dtype=np.float32).reshape(1, -1)
            bias = -(mean @ comp.T).reshape(-1)                      # the centring term, applied after the projection
            dev = torch.device(self.device)
            scale = None
            if self.whiten:
                s = np.sqrt(np.asarray(self.explained_variance_, dtype=np.float32))
                s[s < np.finfo(s.dtype).eps] = np.finfo(s.dtype).eps
                scale = torch.from_numpy(s).to(dev)
            cache = (torch.from_numpy(comp).to(dev), torch.from_numpy(np.ascontiguousarray(bias)).to(dev), scale)
            self._dev_cache = cache
        return cache

    def transform(self, X):
        from sklearn.utils.validation import check_is_fitted
        check_is_fitted(self)
        if self.output not in ("numpy", "torch"):
            raise ValueError("output must be 'numpy' or 'torch'")
        comp, bias, scale = self._device_params()
        if isinstance(X, torch.Tensor):
            x = X.to(comp.device, torch.float32)
        else:
            X = np.asarray(X)
            if X.ndim != 2:
                raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
            x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(comp.device)
        if x.dim() != 2 or x.shape[1] != comp.shape[1]:
            raise ValueError(f"X has {x.shape[-1]} features, but PCA is expecting {comp.shape[1]} features as input")
        y = _project(x.contiguous(), comp, bias)
        if scale is not None:
            y = y / scale
        return y if self.output == "torch" else y.cpu().numpy()

    def __getstate__(self):
        state = super().__getstate__()
        state.pop("_dev_cache", None)
        return state
