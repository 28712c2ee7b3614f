"""1-NN identification parity against the reference's own dependency (scikit-learn KNeighborsClassifier, the real
implementation - facerec_test.py:272,284-285) on seeded galleries."""
import numpy as np
import pytest
import torch
from sklearn import neighbors, preprocessing

import hse_facerec_tf_b200 as hfr

pytestmark = pytest.mark.gpu
# Stated tolerance of the 1-NN parity: the GPU result is the fp64 brute-force answer; sklearn evaluates the expanded form
# in fp64 too, so the two can only differ where the top-2 squared distances agree to fp64 rounding (~1e-13 at |x| ~ 1)
TIE_TOL = 1e-10


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
        clear = (d2[:, 1] - d2[:, 0]) > TIE_TOL       # both sides compute in fp64: only fp64-level ties are excluded
    else:
        clear = np.ones(nq, bool)
    assert clear.mean() > 0.99
    np.testing.assert_array_equal(ind[clear], sk_i[clear])
    assert dist.dtype == np.float64
    # sklearn's expanded form |x|^2+|y|^2-2xy cancels in fp64 (abs error ~1e-16 * |x||y| in d2, i.e. ~1e-8 in a distance
    # near zero); ours accumulates (x-y)^2 in fp64
    np.testing.assert_allclose(dist[clear], sk_d[clear], rtol=1e-9, atol=3e-7 * float(np.abs(g).max()) * np.sqrt(d))
    np.testing.assert_array_equal(clf.predict(q)[clear], sk.predict(q)[clear])
    cert, resc = clf.query_stats()
    assert cert + resc == nq


def brute_force_fp64(g, q, k=1, block=256):
    """fp64 squared distances from the float32 rows, direct differences; ties -> lowest index (stable sort)."""
    g64 = g.astype(np.float64)
    out_d, out_i = [], []
    for a in range(0, len(q), block):
        qq = q[a:a + block].astype(np.float64)
        d2 = np.maximum((qq * qq).sum(1)[:, None] + (g64 * g64).sum(1)[None] - 2.0 * qq @ g64.T, 0.0)
        o = np.argsort(d2, axis=1, kind="stable")[:, :k + 1]
        out_i.append(o)
        out_d.append(np.take_along_axis(d2, o, 1))
    return np.concatenate(out_d), np.concatenate(out_i)


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("n,nq,d", [(40000, 2000, 1024), (30000, 1000, 2048), (150000, 1200, 128)])
def test_hard_queries_random_margins(precision, n, nq, d):
    """Random (not planted) queries: the top-2 margins are far below the bf16 / tf32 rounding of the distance GEMM.  The
    certification bound must send every doubtful query to the exact pass: indices equal the fp64 brute force for EVERY
    query whose fp64 top-2 margin is above fp64 rounding noise - no percentage, no fp32-level tolerance."""
    rs = np.random.RandomState(5 + d)
    g = preprocessing.normalize(rs.randn(n, d).astype(np.float32))
    q = preprocessing.normalize(rs.randn(nq, d).astype(np.float32))
    clf = hfr.KNeighborsClassifier(precision=precision).fit(g, np.arange(len(g)))
    dist, ind = clf.kneighbors(q)
    cert, resc = clf.query_stats()
    bd, bi = brute_force_fp64(g, q)
    clear = (bd[:, 1] - bd[:, 0]) > TIE_TOL
    assert clear.mean() > 0.999
    np.testing.assert_array_equal(ind[clear, 0], bi[clear, 0])
    np.testing.assert_allclose(dist[:, 0] ** 2, bd[:, 0], rtol=0, atol=1e-12)
    assert cert + resc == nq and resc < nq // 4, (cert, resc)       # the exact pass is the exception, not the path
    print(f"hard queries {precision} n={n} d={d}: certified {cert}, re-scored exactly {resc}")


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_certification_bound_covers_the_gemm_error(precision):
    """The bound the certification rests on (knn.cuh): every approximate candidate score the GEMM epilogue wrote differs
    from the exact score |g|^2 - 2 q.g by at most E = c_dot |q| max|g| + c_norm max|g|^2."""
    import ctypes as C
    from hse_facerec_tf_b200._lib import lib
    rs = np.random.RandomState(11)
    n, nq, d = 20000, 512, 1024
    g = (rs.randn(n, d) * rs.uniform(0.2, 3.0, (n, 1))).astype(np.float32)
    q = (rs.randn(nq, d) * rs.uniform(0.2, 3.0, (nq, 1))).astype(np.float32)
    clf = hfr.KNeighborsClassifier(precision=precision).fit(g, np.arange(n))
    clf.kneighbors(q)
    rec = int(lib.hfr_knn_debug_candidates(clf._knn, None, None, nq))
    score = np.zeros((nq, rec), np.float32)
    idx = np.zeros((nq, rec), np.int32)
    assert lib.hfr_knn_debug_candidates(clf._knn, score.ctypes.data, idx.ctypes.data, nq) == rec
    g64, q64 = g.astype(np.float64), q.astype(np.float64)
    gn = (g64 * g64).sum(1)
    u = 2.0 ** -9 if precision == "bf16" else 2.0 ** -10
    c_dot = 2 * (2 * u + u * u + max(d / 2 ** 22, 2.0 ** -12)) + 2.0 ** -23
    c_norm = (d + 4) / 2 ** 24
    worst = 0.0
    for r in range(nq):
        ok = idx[r] >= 0
        j = idx[r][ok]
        exact = gn[j] - 2.0 * (g64[j] @ q64[r])
        E = c_dot * np.sqrt((q64[r] ** 2).sum()) * np.sqrt(gn.max()) + c_norm * gn.max()
        worst = max(worst, float(np.abs(score[r][ok] - exact).max() / E))
    print(f"{precision}: worst |approx - exact| / E = {worst:.3f}")
    assert worst < 1.0


def test_many_duplicates_in_one_bucket_go_through_the_exact_pass():
    """More equal rows than a bucket record holds: the dropped duplicates tie with the kept ones, certification must
    fail and the exact pass must return the lowest index (sklearn's heap keeps the first-seen row)."""
    rs = np.random.RandomState(3)
    g = preprocessing.normalize(rs.randn(3000, 256).astype(np.float32))
    g[100:110] = g[100]                               # ten copies inside one 128-row bucket
    q = g[[105, 2000]]
    for k in (1, 3):
        clf = hfr.KNeighborsClassifier(n_neighbors=k, precision="bf16").fit(g, np.arange(len(g)))
        ind = clf.kneighbors(q, return_distance=False)
        assert ind[0].tolist() == list(range(100, 100 + k)) and ind[1, 0] == 2000
        assert clf.query_stats()[1] >= 1


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


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_sharded_protocol_certifies_globally_and_stays_exact(precision):
    """The row-sharded protocol of include/hfr.h on one GPU (4 handles stand in for 4 ranks, torch.stack for the
    all-gather): hfr_knn_query_partial -> hfr_knn_merge_certify -> (listed queries only) hfr_knn_query_exact ->
    hfr_knn_merge_listed.  Planted queries are certified after the exchange although three of the four shards do not hold
    their neighbour (per-shard certification sent those to the shards' fp64 passes: 7x slower at 2 GPUs); random queries
    and a query with six near-duplicate neighbours are listed, and every result equals the single-gallery answer."""
    import ctypes as C
    from hse_facerec_tf_b200._lib import check, lib, PREC
    n, d = 20000, 512
    g, q, _ = make_problem(n, 400, d, seed=21)         # planted queries (_ = the rows they were planted next to)
    rs = np.random.RandomState(5)
    r0 = next(r for r in range(300, n - 6) if not np.isin(np.arange(r, r + 6), _).any())   # rows no planted query uses
    v = g[r0].copy()
    g[r0:r0 + 6] = v + 1e-5 * rs.randn(6, d).astype(np.float32)
    qr = rs.randn(200, d).astype(np.float32)
    qr /= np.linalg.norm(qr, axis=1, keepdims=True)
    qr[1] = v
    bounds = [0, 3, 5000, 12000, n]
    shards = []
    for a0, b0 in zip(bounds[:-1], bounds[1:]):
        gt = torch.from_numpy(g[a0:b0]).cuda()
        h = C.c_void_p()
        check(lib.hfr_knn_create(0, d, PREC[precision], C.byref(h)))
        check(lib.hfr_knn_set_gallery(h, gt.data_ptr(), b0 - a0, a0, None))
        shards.append((gt, h))

    def sharded_query(queries, k):
        qt = torch.from_numpy(queries).cuda()
        nq = len(queries)
        parts = []
        for _, h in shards:
            o = torch.empty((nq, k + 1, 2), dtype=torch.int64, device="cuda")
            check(lib.hfr_knn_query_partial(h, qt.data_ptr(), nq, k, o.data_ptr(), None))
            parts.append(o)
        parts = torch.stack(parts).contiguous()
        out = torch.empty((nq, k, 2), dtype=torch.int64, device="cuda")
        unc = torch.empty((nq + 1,), dtype=torch.int32, device="cuda")
        check(lib.hfr_knn_merge_certify(parts.data_ptr(), len(shards), nq, k, out.data_ptr(), unc[1:].data_ptr(),
                                        unc.data_ptr(), 0, None))
        n_unc = int(unc[0].item())
        if n_unc:
            locs = []
            for _, h in shards:
                loc = torch.empty((nq, k, 2), dtype=torch.int64, device="cuda")
                check(lib.hfr_knn_query_exact(h, qt.data_ptr(), nq, k, unc[1:].data_ptr(), unc.data_ptr(), loc.data_ptr(), None))
                locs.append(loc)
            parts2 = torch.stack(locs).contiguous()
            check(lib.hfr_knn_merge_listed(parts2.data_ptr(), len(shards), nq, k, unc[1:].data_ptr(), unc.data_ptr(),
                                           out.data_ptr(), 0, None))
        torch.cuda.synchronize()
        rec = out.cpu().numpy()
        listed = set(unc[1:1 + n_unc].cpu().numpy().tolist())
        return np.ascontiguousarray(rec[:, :, 0]).view(np.float64), rec[:, :, 1], listed

    for k in (1, 3):
        whole = hfr.KNeighborsClassifier(n_neighbors=k, precision=precision).fit(g, np.arange(n))
        d2, ind, listed = sharded_query(q, k)
        d_ref, i_ref = whole.kneighbors(q)
        np.testing.assert_array_equal(ind, i_ref)
        np.testing.assert_array_equal(np.sqrt(d2), d_ref)
        if k == 1:
            assert not listed, listed                    # planted neighbours: certified globally
        d2, ind, listed = sharded_query(qr, k)
        d_ref, i_ref = whole.kneighbors(qr)
        np.testing.assert_array_equal(ind, i_ref)
        np.testing.assert_array_equal(np.sqrt(d2), d_ref)
        assert 1 in listed                               # six near-duplicates, four re-scored: cannot be certified
        assert r0 <= ind[1, 0] < r0 + 6
    for _, h in shards:
        lib.hfr_knn_free(h)


def test_merge_of_shards_equals_single_gallery():
    """Gallery row-sharded over 4 handles on one GPU + hfr_knn_merge == one handle over the whole gallery (k = 1 and 3,
    duplicates across shards, a shard smaller than k)."""
    import ctypes as C
    from hse_facerec_tf_b200._lib import check, lib
    g, q, _ = make_problem(10000, 500, 512, seed=9)
    g[9000] = g[7]                                    # duplicate across shards: lowest global index wins
    q[0] = g[7]
    qt = torch.from_numpy(q).cuda()
    bounds = [0, 2, 2500, 7000, 10000]                # first shard has 2 rows (< k = 3)
    for k in (1, 3):
        whole = hfr.KNeighborsClassifier(n_neighbors=k, precision="bf16").fit(g, np.arange(len(g)))
        d_ref, i_ref = whole.kneighbors(q)
        outs, keep = [], []
        for a0, b0 in zip(bounds[:-1], bounds[1:]):
            gt = torch.from_numpy(g[a0:b0]).cuda()
            h = C.c_void_p()
            check(lib.hfr_knn_create(0, 512, 2, C.byref(h)))
            check(lib.hfr_knn_set_gallery(h, gt.data_ptr(), b0 - a0, a0, None))
            o = torch.empty((len(q), k, 2), dtype=torch.int64, device="cuda")
            check(lib.hfr_knn_query(h, qt.data_ptr(), len(q), k, o.data_ptr(), None))
            outs.append(o)
            keep.append((gt, h))
        parts = torch.stack(outs).contiguous()
        merged = torch.empty((len(q), k, 2), dtype=torch.int64, device="cuda")
        check(lib.hfr_knn_merge(parts.data_ptr(), len(outs), len(q), k, merged.data_ptr(), 0, None))
        torch.cuda.synchronize()
        rec = merged.cpu().numpy()
        np.testing.assert_array_equal(rec[:, :, 1], i_ref)
        np.testing.assert_array_equal(np.sqrt(np.ascontiguousarray(rec[:, :, 0]).view(np.float64)), d_ref)
        assert rec[0, 0, 1] == 7
        if k == 3:
            first = outs[0].cpu().numpy()
            assert (first[:, 2, 1] == -1).all() and np.isinf(np.ascontiguousarray(first[:, 2, 0]).view(np.float64)).all()
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
    clear = (np.diff(dk, axis=1) > TIE_TOL).all(axis=1)          # no fp64-level ties among the first k+1 exact distances
    assert clear.mean() > 0.99
    np.testing.assert_array_equal(ind[clear], sk_i[clear])
    np.testing.assert_allclose(dist[clear], sk_d[clear], rtol=1e-9, atol=1e-7)
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
