"""1-NN identification parity against the reference's own dependency (scikit-learn KNeighborsClassifier, the real
implementation - facerec_test.py:272,284-285) on seeded galleries."""
import numpy as np
import pytest
import torch
from sklearn import neighbors, preprocessing

import hse_facerec_tf_b200 as hfr

pytestmark = pytest.mark.gpu


def make_problem(n, nq, d, seed=0, normalised=True, sigma=0.05):
    rs = np.random.RandomState(seed)
    g = rs.randn(n, d).astype(np.float32)
    if normalised:
        g = preprocessing.normalize(g)
    else:
        g *= rs.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)   # exercises the ||g||^2 term
    pick = np.random.RandomState(seed + 1).randint(0, n, nq)
    q = g[pick] + sigma * np.random.RandomState(seed + 2).randn(nq, d).astype(np.float32)
    if normalised:
        q = preprocessing.normalize(q)
    return g, q.astype(np.float32), pick


def sk_margins(g, q):
    nn = neighbors.NearestNeighbors(n_neighbors=2, algorithm="brute").fit(g)
    dist, ind = nn.kneighbors(q)
    return dist, ind


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("n,nq,d,normalised", [(5000, 300, 1024, True), (20000, 1000, 2048, True),
                                               (3000, 257, 128, False), (70000, 700, 1024, False), (1, 5, 64, True),
                                               # CTA-pair tiles with an odd number of 128-query blocks (the last pair's
                                               # second CTA is entirely out of range); PCA-sized rows (D = 16, 128)
                                               (40000, 1100, 1024, True), (30000, 130, 16, False), (2000, 50, 16, True),
                                               (50000, 900, 128, False)])
def test_kneighbors_matches_sklearn(precision, n, nq, d, normalised):
    g, q, pick = make_problem(n, nq, d, seed=n % 97, normalised=normalised)
    y = np.arange(n) % 1000
    clf = hfr.KNeighborsClassifier(n_neighbors=1, p=2, precision=precision).fit(g, y)
    dist, ind = clf.kneighbors(q)
    sk = neighbors.KNeighborsClassifier(n_neighbors=1, p=2).fit(g, y)
    sk_d, sk_i = sk.kneighbors(q)
    if n >= 2:
        d2, _ = sk_margins(g, q)
        clear = (d2[:, 1] - d2[:, 0]) > 1e-6          # stated tolerance: fp64 re-rank => only exact ties are excluded
    else:
        clear = np.ones(nq, bool)
    assert clear.mean() > 0.99
    np.testing.assert_array_equal(ind[clear], sk_i[clear])
    np.testing.assert_allclose(dist[clear], sk_d[clear], rtol=1e-4, atol=2e-4)
    np.testing.assert_array_equal(clf.predict(q)[clear], sk.predict(q)[clear])


def test_hard_queries_random_margins():
    """Random (not planted) queries: top-2 margins are tiny; the fp32/fp64 re-rank of the bf16 candidates must still
    agree with sklearn wherever the margin exceeds 1e-5."""
    rs = np.random.RandomState(5)
    g = preprocessing.normalize(rs.randn(40000, 1024).astype(np.float32))
    q = preprocessing.normalize(rs.randn(2000, 1024).astype(np.float32))
    clf = hfr.KNeighborsClassifier(precision="bf16").fit(g, np.arange(len(g)))
    ind = clf.kneighbors(q, return_distance=False)[:, 0]
    d2, i2 = sk_margins(g, q)
    clear = (d2[:, 1] - d2[:, 0]) > 1e-5
    agree = (ind == i2[:, 0])
    assert agree[clear].mean() > 0.995, agree[clear].mean()   # candidates come from a bf16 top-2 per 16k-row split


def test_exact_duplicates_tie_to_lowest_index():
    rs = np.random.RandomState(2)
    g = rs.randn(1200, 256).astype(np.float32)
    g[500] = g[10]
    g[900] = g[10]
    clf = hfr.KNeighborsClassifier(precision="tf32").fit(g, np.arange(1200))
    assert clf.predict(g[10:11])[0] == 10


def test_sklearn_protocol(tmp_path):
    from sklearn import model_selection
    from sklearn.base import clone
    from sklearn.decomposition import PCA
    from sklearn.pipeline import Pipeline
    g, q, pick = make_problem(2000, 10, 256, seed=3)
    y = np.repeat(np.arange(500), 4)
    X = preprocessing.normalize(g + 0.0)
    clf = hfr.KNeighborsClassifier(n_neighbors=1, p=2)
    assert clone(clf).get_params()["n_neighbors"] == 1
    # the reference's evaluation protocol, facerec_test.py:200-207
    sss = model_selection.StratifiedShuffleSplit(n_splits=1, test_size=0.5, random_state=0)
    ours = model_selection.cross_validate(clf, X, y, scoring="accuracy", cv=sss)["test_score"]
    ref = model_selection.cross_validate(neighbors.KNeighborsClassifier(n_neighbors=1, p=2), X, y, scoring="accuracy",
                                         cv=sss)["test_score"]
    np.testing.assert_allclose(ours, ref)
    pipe = Pipeline([("pca", PCA(n_components=20, random_state=0)), ("classifier", hfr.KNeighborsClassifier(1, 2))])
    ref_pipe = Pipeline([("pca", PCA(n_components=20, random_state=0)),
                         ("classifier", neighbors.KNeighborsClassifier(n_neighbors=1, p=2))])
    assert (pipe.fit(X[::2], y[::2]).predict(X[1::2]) == ref_pipe.fit(X[::2], y[::2]).predict(X[1::2])).mean() > 0.99
    with pytest.raises(Exception):
        hfr.KNeighborsClassifier().predict(X)


def test_merge_of_shards_equals_single_gallery():
    """Gallery row-sharded over 4 handles on one GPU + hfr_knn_merge == one handle over the whole gallery."""
    import ctypes as C
    from hse_facerec_tf_b200._lib import check, lib
    g, q, _ = make_problem(10000, 500, 512, seed=9)
    whole = hfr.KNeighborsClassifier(precision="bf16").fit(g, np.arange(len(g)))
    d_ref, i_ref = whole.kneighbors(q)
    parts = np.array_split(np.arange(len(g)), 4)
    d_all, i_all = [], []
    qt = torch.from_numpy(q).cuda()
    keep = []
    for p in parts:
        gt = torch.from_numpy(g[p]).cuda()
        h = C.c_void_p()
        check(lib.hfr_knn_create(0, 512, 2, C.byref(h)))
        check(lib.hfr_knn_set_gallery(h, gt.data_ptr(), len(p), int(p[0]), None))
        d = torch.empty(len(q), device="cuda")
        i = torch.empty(len(q), dtype=torch.int64, device="cuda")
        check(lib.hfr_knn_query(h, qt.data_ptr(), len(q), d.data_ptr(), i.data_ptr(), None))
        d_all.append(d)
        i_all.append(i)
        keep.append((gt, h))
    torch.cuda.synchronize()
    D, I = torch.stack(d_all).contiguous(), torch.stack(i_all).contiguous()
    bd = torch.empty(len(q), device="cuda")
    bi = torch.empty(len(q), dtype=torch.int64, device="cuda")
    check(lib.hfr_knn_merge(D.data_ptr(), I.data_ptr(), 4, len(q), bd.data_ptr(), bi.data_ptr(), 0, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(bi.cpu().numpy(), i_ref[:, 0])
    for gt, h in keep:
        lib.hfr_knn_free(h)


def test_empty_and_tiny_inputs():
    g = np.random.RandomState(0).randn(3, 64).astype(np.float32)
    clf = hfr.KNeighborsClassifier(precision="tf32").fit(g, np.array([7, 8, 9]))
    assert clf.predict(np.zeros((0, 64), np.float32)).shape == (0,)
    d, i = clf.kneighbors(g[1:2])
    assert i[0, 0] == 1 and d[0, 0] < 1e-6
    with pytest.raises(ValueError):
        clf.predict(np.zeros((2, 32), np.float32))          # wrong feature count, as sklearn
    with pytest.raises(ValueError):
        hfr.KNeighborsClassifier().fit(np.zeros((0, 64), np.float32), np.zeros(0))
    with pytest.raises(ValueError):
        hfr.KNeighborsClassifier(n_neighbors=7).fit(g, np.array([7, 8, 9]))      # the GPU path stops at k = 4
    assert hfr.KNeighborsClassifier(n_neighbors=3, precision="tf32").fit(g, np.array([7, 8, 9])).predict(g[:1])[0] == 7


def test_extract_then_identify_end_to_end(age_gender_pb, golden_dir):
    """BASELINE config 5 as a parity case: uint8 crops -> GPU embeddings (L2-normalised in the same call) -> GPU 1-NN,
    against oracle embeddings + scikit-learn (the reference's own pipeline, facerec_test.py:394-442)."""
    from oracle.tfnet import GraphOracle, preprocess_rgb_u8
    from tests.helpers import smooth_images
    crops = np.load(f"{golden_dir}/face_crops_u8.npz")["c192"]
    imgs = np.concatenate([crops, smooth_images(12, 192, 21)])
    (ref,) = GraphOracle(age_gender_pb).run(["global_pooling/Mean:0"], {"input_1:0": preprocess_rgb_u8(imgs)})
    ref = preprocessing.normalize(ref)
    rs = np.random.RandomState(8)
    distract = preprocessing.normalize(np.abs(rs.randn(20000, 1024)).astype(np.float32))   # embeddings are non-negative
    gallery = np.concatenate([ref, distract]).astype(np.float32)
    y = np.arange(len(gallery)) % 5000
    sk_pred = neighbors.KNeighborsClassifier(n_neighbors=1, p=2).fit(gallery, y).predict(ref)
    for precision in ("tf32", "bf16"):
        tfi = hfr.TensorFlowInference(age_gender_pb, "input_1:0", "global_pooling/Mean:0", precision=precision, input_hw=192)
        emb = tfi.extract_batch(torch.from_numpy(imgs).cuda(), l2norm=True)
        clf = hfr.KNeighborsClassifier(n_neighbors=1, p=2, precision=precision).fit(gallery, y)
        pred = clf.predict(emb)                      # CUDA tensor in, labels out
        np.testing.assert_array_equal(pred, sk_pred)
        np.testing.assert_array_equal(pred, y[: len(imgs)])


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("k", [2, 3, 4])
@pytest.mark.parametrize("n,nq,d", [(5000, 300, 1024), (40000, 700, 128), (3, 20, 64)])
def test_k_nearest_neighbours_match_sklearn(precision, k, n, nq, d):
    """KNeighborsClassifier(n_neighbors=3, p=2) of the reference's list (facerec_test.py:274-275): neighbour lists equal
    sklearn's wherever consecutive exact distances differ by more than the stated tolerance, and the uniform vote -
    ties to the smallest class - gives the same labels."""
    if k > n:
        pytest.skip("k > gallery size is an error, covered below")
    g, q, _ = make_problem(n, nq, d, seed=7 + k, normalised=True)
    y = (np.arange(n) * 7) % 23                                  # few classes: votes do collide
    clf = hfr.KNeighborsClassifier(n_neighbors=k, p=2, precision=precision).fit(g, y)
    sk = neighbors.KNeighborsClassifier(n_neighbors=k, p=2).fit(g, y)
    sk_d, sk_i = sk.kneighbors(q)
    dist, ind = clf.kneighbors(q)
    assert dist.shape == ind.shape == (nq, k)
    kk = min(k + 1, n)
    dk = neighbors.NearestNeighbors(n_neighbors=kk, algorithm="brute").fit(g).kneighbors(q)[0] ** 2
    clear = (np.diff(dk, axis=1) > 1e-6).all(axis=1)             # no (near-)ties among the first k+1 exact distances
    assert clear.mean() > 0.95
    np.testing.assert_array_equal(ind[clear], sk_i[clear])
    np.testing.assert_allclose(dist[clear], sk_d[clear], rtol=1e-4, atol=2e-4)
    np.testing.assert_array_equal(clf.predict(q)[clear], sk.predict(q)[clear])
    # a per-call override, as sklearn allows
    np.testing.assert_array_equal(clf.kneighbors(q, n_neighbors=1, return_distance=False)[clear][:, 0], sk_i[clear][:, 0])


def test_k_neighbours_argument_errors():
    g, q, _ = make_problem(50, 5, 64, seed=1)
    with pytest.raises(ValueError):
        hfr.KNeighborsClassifier(n_neighbors=5).fit(g, np.arange(50))      # beyond the GPU path's k <= 4
    clf = hfr.KNeighborsClassifier(n_neighbors=3).fit(g[:2], np.arange(2))
    with pytest.raises(ValueError):                                          # sklearn: n_neighbors <= n_samples_fit
        clf.kneighbors(q)
    np.testing.assert_array_equal(clf.kneighbors(q, n_neighbors=2, return_distance=False).shape, (5, 2))


@pytest.mark.parametrize("n_components,k", [(16, 1), (16, 3), (128, 1)])
def test_pca_knn_pipeline_matches_sklearn(n_components, k):
    """The reference's '1-NN+PCA' / '3-NN+PCA' entries (facerec_test.py:271-274, 417-422) with both steps on the GPU
    against the same Pipeline in scikit-learn: the projection to fp32 rounding, the predictions wherever the exact
    neighbour distances in the projected space are not (near-)tied."""
    from sklearn.decomposition import PCA as SkPCA
    from sklearn.pipeline import Pipeline
    rs = np.random.RandomState(n_components + k)
    n_ids, per_id, d = 120, 8, 1024
    centres = preprocessing.normalize(rs.randn(n_ids, d))
    X = preprocessing.normalize(np.repeat(centres, per_id, 0) + 0.8 * rs.randn(n_ids * per_id, d) / np.sqrt(d)).astype(np.float32)
    y = np.repeat(np.arange(n_ids), per_id)
    test = rs.rand(len(X)) < 0.4
    Xtr, ytr, Xte = X[~test], y[~test], X[test]
    ours = Pipeline([("pca", hfr.PCA(n_components, svd_solver="full")), ("classifier", hfr.KNeighborsClassifier(n_neighbors=k, p=2, precision="tf32"))])
    ref = Pipeline([("pca", SkPCA(n_components, svd_solver="full")), ("classifier", neighbors.KNeighborsClassifier(n_neighbors=k, p=2))])
    ours.fit(Xtr, ytr)
    ref.fit(Xtr, ytr)
    Z, Zr = ours.named_steps["pca"].transform(Xte), ref.named_steps["pca"].transform(Xte)
    assert Z.dtype == np.float32
    np.testing.assert_allclose(Z, Zr, rtol=1e-4, atol=2e-5)
    dk = neighbors.NearestNeighbors(n_neighbors=k + 1, algorithm="brute").fit(ref.named_steps["pca"].transform(Xtr)).kneighbors(Zr)[0] ** 2
    # the two projections differ by fp32 summation order (~3e-6 per coordinate), i.e. up to ~1e-5 in a squared distance
    clear = (np.diff(dk, axis=1) > 1e-4).all(axis=1)
    assert clear.mean() > 0.8
    np.testing.assert_array_equal(ours.predict(Xte)[clear], ref.predict(Xte)[clear])
    assert abs(ours.score(Xte, y[test]) - ref.score(Xte, y[test])) < 0.02
    # device-resident hand-over between the two steps
    pca_t = hfr.PCA(n_components, svd_solver="full", output="torch").fit(Xtr)
    zt = pca_t.transform(torch.from_numpy(Xte).cuda())
    assert zt.is_cuda and np.array_equal(zt.cpu().numpy(), Z)
