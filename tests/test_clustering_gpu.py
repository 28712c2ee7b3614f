"""Distance matrices of the clustering scripts (scope row 8f-3) against what the reference calls: sklearn's
pairwise_distances (facial_clustering_test.py:396) and the per-pair feature_distance of process_photos.py:46-56."""
import numpy as np
import pytest
import torch
from sklearn import preprocessing
from sklearn.metrics import pairwise_distances

import hse_facerec_tf_b200 as hfr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,d", [(300, 300, 1024), (65, 130, 2048), (1, 7, 16), (517, 64, 100)])
def test_pairwise_distances_match_sklearn(n, m, d):
    rs = np.random.RandomState(n + d)
    X = preprocessing.normalize(rs.randn(n, d)).astype(np.float32)
    Y = preprocessing.normalize(rs.randn(m, d)).astype(np.float32)
    X[-1] = X[0]                                                   # an exact duplicate: distance exactly 0
    np.testing.assert_allclose(hfr.pairwise_distances(X, Y), pairwise_distances(X, Y), rtol=1e-5, atol=1e-6)
    D = hfr.pairwise_distances(X)
    ref = pairwise_distances(X)
    np.testing.assert_allclose(D, ref, rtol=1e-5, atol=1e-6)
    assert (np.diag(D) == 0).all() and D[0, -1] == 0 and np.array_equal(D, D.T)
    t = hfr.pairwise_distances(torch.from_numpy(X).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), D)
    with pytest.raises(ValueError):
        hfr.pairwise_distances(X, Y[:, :-1])


def test_album_distance_matrix_matches_process_photos():
    rs = np.random.RandomState(3)
    n, d = 97, 1024
    f = preprocessing.normalize(rs.randn(n, d)).astype(np.float32)
    years = rs.randint(2005, 2019, n)
    born = years - rs.randint(1, 70, n)                             # apparent age at photo time >= 1

    # process_photos.py:46-56 restated with broadcasting: euclidean distance between the embeddings plus a tenth of
    # (a_i - a_j)^2 / (a_i + a_j), a = apparent age of each face in the year of the later of the two photos; clipped at 0
    later = np.maximum(years[:, None], years[None, :]).astype(np.float64)
    age_i, age_j = later - born[:, None], later - born[None, :]
    feat = np.sqrt(((f[:, None, :].astype(np.float64) - f[None, :, :].astype(np.float64)) ** 2).sum(axis=2))
    ref = np.clip(feat + 0.1 * (age_i - age_j) ** 2 / (age_i + age_j), 0, None)
    got = hfr.album_distance_matrix(f, years, born)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError):
        hfr.album_distance_matrix(f, years[:-1], born)
