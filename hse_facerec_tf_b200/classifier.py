"""Drop-in for the reference's classifier: KNeighborsClassifier(n_neighbors=1, p=2) (facerec_test.py:272,422) - and the
n_neighbors=3 entries of the same list (facerec_test.py:274-275; uniform weights, 1 <= n_neighbors <= 4 here) - used
through the sklearn estimator protocol by classifier_tester / cross_validate (facerec_test.py:200-207) and by direct
fit/predict (facerec_test.py:282-287,434-442); also valid as the last step of a Pipeline (facerec_test.py:271,421).

fit stores the gallery on the GPU (row-sharded across ranks when a torch.distributed process group is given);
predict runs the fused distance-GEMM + argmin kernel and, for a sharded gallery, merges the per-rank
(distance, index) pairs after one all-gather over NCCL.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from sklearn.base import BaseEstimator, ClassifierMixin

from . import _lib
from ._lib import check, lib
from .model import _stream_ptr


def uniform_vote(enc: np.ndarray, n_classes: int) -> np.ndarray:
    """enc: [nq, k] class indices (into the sorted classes_) of every query's neighbours -> [nq] winning class index under
    uniform weights: most votes, a tie going to the smallest class index - sklearn's predict takes the argmax of the
    class-probability row (neighbors/_classification.py), and argmax returns the first maximum."""
    votes = (enc[:, :, None] == enc[:, None, :]).sum(axis=2)          # votes[i, j] = neighbours sharing neighbour j's class
    key = votes.astype(np.int64) * (n_classes + 1) - enc              # more votes first, then the smaller class
    return enc[np.arange(len(enc)), key.argmax(axis=1)]


class KNeighborsClassifier(ClassifierMixin, BaseEstimator):
    MAX_NEIGHBORS = 4

    def __init__(self, n_neighbors=1, p=2, *, device="cuda:0", precision="bf16", process_group=None, sharded=False):
        self.n_neighbors = n_neighbors
        self.p = p
        self.device = device
        self.precision = precision
        self.process_group = process_group
        self.sharded = sharded

    # -- helpers --------------------------------------------------------------------------------------
    def _dev(self):
        return torch.device(self.device)

    def _to_dev(self, X):
        if isinstance(X, torch.Tensor):
            t = X.to(self._dev(), torch.float32)
        else:
            X = np.asarray(X)
            if X.ndim != 2:
                raise ValueError(f"Expected 2D array, got {X.ndim}D array instead")
            t = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(self._dev())
        return t.contiguous()

    def _free(self):
        h = getattr(self, "_knn", None)
        if h:
            lib.hfr_knn_free(h)
            self._knn = None

    def __del__(self):
        self._free()

    # -- sklearn protocol -------------------------------------------------------------------------------
    def fit(self, X, y):
        """X: (N, D) gallery embeddings.  With sharded=True each rank passes ITS rows of the gallery (contiguous
        blocks in rank order) and the labels of the whole gallery are all-gathered."""
        if self.p != 2 or not isinstance(self.n_neighbors, (int, np.integer)) or not 1 <= self.n_neighbors <= self.MAX_NEIGHBORS:
            raise ValueError(f"the GPU path covers p=2 and 1 <= n_neighbors <= {self.MAX_NEIGHBORS} "
                             "(the reference uses n_neighbors=1 and 3)")
        if not torch.cuda.is_available():
            raise _lib.HfrError("no CUDA device available; this classifier has no CPU fallback")
        self._free()
        g = self._to_dev(X)
        y = np.asarray(y.cpu() if isinstance(y, torch.Tensor) else y)
        if g.shape[0] != y.shape[0]:
            raise ValueError("X and y have inconsistent numbers of samples")
        if g.shape[0] == 0:
            raise ValueError("Found array with 0 sample(s)")
        self._pad = (-g.shape[1]) % 8  # TMA rows must be multiples of 16 bytes: zero-pad odd dimensions (e.g. PCA)
        if self._pad:
            g = torch.nn.functional.pad(g, (0, self._pad))
        self.n_features_in_ = int(g.shape[1] - self._pad)
        offset = 0
        if self.sharded:
            from .parallel import shard_layout
            offset, y, _ = shard_layout(g.shape[0], y, self.process_group)
        self.classes_, self._y = np.unique(y, return_inverse=True)
        self._labels = y
        self._gallery = g  # keeps the fp32 rows alive: the kernel re-ranks its candidates against them
        h = C.c_void_p()
        dev = self._dev().index or 0
        check(lib.hfr_knn_create(dev, g.shape[1], _lib.PREC[self.precision], C.byref(h)))
        self._knn = h
        check(lib.hfr_knn_set_gallery(h, g.data_ptr(), g.shape[0], offset, _stream_ptr(g.device)))
        return self

    def kneighbors(self, X, n_neighbors=None, return_distance=True, local_queries=False, total_queries=None):
        """sklearn's kneighbors.  local_queries=True (sharded gallery only): X holds THIS rank's block of the queries
        (contiguous blocks in rank order, e.g. the embeddings it extracted from its slice of the batch); the blocks
        are all-gathered over NCCL and every rank returns the neighbours of the whole query set.  total_queries: the
        size of the whole set when the blocks follow parallel.shard_rows (saves the exchange of the block sizes)."""
        if getattr(self, "_knn", None) is None:
            from sklearn.exceptions import NotFittedError
            raise NotFittedError("This KNeighborsClassifier instance is not fitted yet.")
        k = self.n_neighbors if n_neighbors is None else n_neighbors
        if not isinstance(k, (int, np.integer)) or not 1 <= k <= self.MAX_NEIGHBORS:
            raise ValueError(f"n_neighbors must be an integer in 1..{self.MAX_NEIGHBORS}, got {k!r}")
        if k > len(self._labels):
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = "
                             f"{len(self._labels)}, n_samples = {len(X)}")
        q = self._to_dev(X)
        if q.shape[1] != self.n_features_in_:
            raise ValueError(f"X has {q.shape[1]} features, but KNeighborsClassifier is expecting "
                             f"{self.n_features_in_} features as input")
        if self._pad:
            q = torch.nn.functional.pad(q, (0, self._pad))
        if local_queries and self.sharded:
            from .parallel import gather_rows
            q = gather_rows(q, self.process_group, total_queries)
        nq = q.shape[0]
        if nq == 0:
            empty = np.zeros((0, k), np.int64)
            return (np.zeros((0, k), np.float64), empty) if return_distance else empty
        dev = q.device.index or 0
        stream = _stream_ptr(q.device)
        # hfr_neighbor records {double dist2; int64 index} as one int64 tensor [nq, k, 2]
        out = torch.empty((nq, k, 2), dtype=torch.int64, device=q.device)
        if not self.sharded:
            check(lib.hfr_knn_query(self._knn, q.data_ptr(), nq, int(k), out.data_ptr(), stream))
            self._sharded_unc = None
        else:
            # Row-sharded gallery (include/hfr.h, "Row-sharded gallery, exact and shard-independent"): every shard
            # proposes its re-scored candidates plus a lower bound for everything it did not re-score; after ONE
            # all-gather the global answer is certified against the smallest bound.  Only queries that fail it (the
            # same list on every rank) go through the shards' fp64 passes and a second exchange.
            from .parallel import gather_neighbors
            part = torch.empty((nq, k + 1, 2), dtype=torch.int64, device=q.device)
            check(lib.hfr_knn_query_partial(self._knn, q.data_ptr(), nq, int(k), part.data_ptr(), stream))
            parts = gather_neighbors(part, self.process_group)
            unc = torch.empty((nq + 1,), dtype=torch.int32, device=q.device)     # [0]: count, [1:]: the list
            check(lib.hfr_knn_merge_certify(parts.data_ptr(), parts.shape[0], nq, int(k), out.data_ptr(),
                                            unc[1:].data_ptr(), unc.data_ptr(), dev, stream))
            n_unc = int(unc[:1].cpu()[0])
            if n_unc:
                loc = torch.empty((nq, k, 2), dtype=torch.int64, device=q.device)
                check(lib.hfr_knn_query_exact(self._knn, q.data_ptr(), nq, int(k), unc[1:].data_ptr(), unc.data_ptr(),
                                              loc.data_ptr(), stream))
                parts2 = gather_neighbors(loc, self.process_group)
                check(lib.hfr_knn_merge_listed(parts2.data_ptr(), parts2.shape[0], nq, int(k), unc[1:].data_ptr(),
                                               unc.data_ptr(), out.data_ptr(), dev, stream))
            self._sharded_unc = (nq, n_unc)
        self._last_out = out
        rec = out.cpu().numpy()
        ind = np.ascontiguousarray(rec[:, :, 1])
        if return_distance:
            d2 = np.ascontiguousarray(rec[:, :, 0]).view(np.float64)   # fp64, accumulated in fp64 on the GPU
            return np.sqrt(d2), ind
        return ind

    def query_stats(self):
        """(queries certified by the rounding bound, queries re-scored exactly against the whole shard) of the last
        kneighbors call on this rank."""
        if getattr(self, "_sharded_unc", None) is not None:     # sharded: certified globally, after the exchange
            nq, n_unc = self._sharded_unc
            return nq - n_unc, n_unc
        a, b = C.c_int64(), C.c_int64()
        check(lib.hfr_knn_stats(self._knn, C.byref(a), C.byref(b)))
        return a.value, b.value

    def predict(self, X):
        """Uniform-weight vote over the n_neighbors nearest rows; a tie goes to the smallest class (sklearn takes the
        argmax of the class-probability row, classes_ being sorted).  n_neighbors=1 is a plain label lookup."""
        ind = self.kneighbors(X, return_distance=False)
        if ind.shape[1] == 1:
            return self._labels[ind[:, 0]]
        return self.classes_[uniform_vote(self._y[ind], len(self.classes_))]
