"""N>1 host logic on CPU: two gloo processes shard a gallery, score their shard with the oracle (scikit-learn), exchange
(distance, index) pairs through hse_facerec_tf_b200.parallel and must reproduce the single-gallery 1-NN result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from sklearn import neighbors

from hse_facerec_tf_b200 import parallel
from tests.helpers import merge_pairs_reference


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rs = np.random.RandomState(0)
    g = rs.randn(1001, 32).astype(np.float32)
    g[700] = g[3]                        # duplicate rows across the shard boundary: tie -> lowest global index
    q = np.concatenate([g[[3, 500, 1000]], rs.randn(40, 32).astype(np.float32)])
    y = (np.arange(1001) * 7) % 113
    return g, q, y


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g, q, y = _problem()
        a, b = parallel.shard_rows(len(g), ws, rank)
        offset, labels, counts = parallel.shard_layout(b - a, y[a:b])
        assert offset == a and counts == [parallel.shard_rows(len(g), ws, r)[1] - parallel.shard_rows(len(g), ws, r)[0]
                                          for r in range(ws)]
        np.testing.assert_array_equal(labels, y)
        nn = neighbors.NearestNeighbors(n_neighbors=1, algorithm="brute").fit(g[a:b])
        d, i = nn.kneighbors(q)
        # the packed hfr_neighbor records {double dist2; int64 index} the GPU path exchanges, [nq, k=1, 2] int64
        rec = np.zeros((len(q), 1, 2), np.int64)
        rec[:, 0, 0] = (d[:, 0] ** 2).astype(np.float64).view(np.int64)
        rec[:, 0, 1] = i[:, 0].astype(np.int64) + offset
        parts = parallel.gather_neighbors(torch.from_numpy(rec)).numpy()
        assert parts.shape == (ws, len(q), 1, 2)
        _, best = merge_pairs_reference(np.ascontiguousarray(parts[:, :, 0, 0]).view(np.float64), parts[:, :, 0, 1])
        emb = parallel.gather_rows(torch.from_numpy(g[a:b]))          # embeddings all-gather with ragged blocks
        assert torch.equal(emb, torch.from_numpy(g))
        mine = parallel.split_batch(np.arange(10))
        assert list(mine) == list(range(*parallel.shard_rows(10, ws, rank)))
        if rank == 0:
            out.put(best)
    finally:
        dist.destroy_process_group()


def test_sharded_gallery_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    best = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g, q, y = _problem()
    ref = neighbors.KNeighborsClassifier(n_neighbors=1, p=2).fit(g, y)
    ref_i = ref.kneighbors(q, return_distance=False)[:, 0]
    assert best[0] == 3                      # duplicate of row 3 lives in shard 1 as row 700: lowest index wins
    np.testing.assert_array_equal(best[1:], ref_i[1:])
    np.testing.assert_array_equal(y[best], ref.predict(q))


def test_shard_rows_partition():
    for n in (0, 1, 7, 1000, 1001):
        for ws in (1, 2, 3, 8):
            blocks = [parallel.shard_rows(n, ws, r) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_uniform_vote_matches_sklearn_predict():
    """The k-NN vote the classifier applies to the GPU's neighbour lists, fed here with sklearn's own lists: equal to
    sklearn's predict for k = 2, 3, 4, including three-way ties (which go to the smallest class, not the nearest)."""
    from hse_facerec_tf_b200.classifier import uniform_vote
    rs = np.random.RandomState(5)
    g = rs.randn(400, 6).astype(np.float32)
    q = rs.randn(300, 6).astype(np.float32)
    y = rs.randint(0, 9, 400) * 5 - 7                       # non-contiguous, negative labels
    classes, enc_all = np.unique(y, return_inverse=True)
    for k in (2, 3, 4):
        sk = neighbors.KNeighborsClassifier(n_neighbors=k, p=2).fit(g, y)
        ind = sk.kneighbors(q, return_distance=False)
        pred = classes[uniform_vote(enc_all[ind], len(classes))]
        np.testing.assert_array_equal(pred, sk.predict(q))
        assert (np.array([len(set(r)) for r in enc_all[ind]]) == k).any()      # all-different rows did occur


def test_pca_wrapper_logic_matches_sklearn_transform(monkeypatch):
    """hse_facerec_tf_b200.PCA = sklearn's fit + a GPU projection; with the projection replaced by a host matmul the
    wrapper (centring folded into a bias, whitening, clone/get_params, pickling) must reproduce sklearn's transform."""
    import pickle
    from sklearn.base import clone
    from sklearn.decomposition import PCA as SkPCA
    import hse_facerec_tf_b200.decomposition as dec
    monkeypatch.setattr(dec, "_project", lambda x, comp, bias: x @ comp.T + bias)
    rs = np.random.RandomState(2)
    X = (rs.randn(300, 64) @ rs.randn(64, 64)).astype(np.float32) + 3.0
    Q = (rs.randn(40, 64) @ rs.randn(64, 64)).astype(np.float32) + 3.0
    for whiten in (False, True):
        ours = dec.PCA(16, whiten=whiten, svd_solver="full", device="cpu").fit(X)
        ref = SkPCA(16, whiten=whiten, svd_solver="full").fit(X)
        got = ours.transform(Q)
        assert got.dtype == np.float32 and got.shape == (40, 16)
        np.testing.assert_allclose(got, ref.transform(Q), rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(ours.fit_transform(X), ref.fit_transform(X), rtol=2e-3, atol=2e-3)
        twin = clone(ours)
        assert twin.get_params()["whiten"] == whiten and twin.get_params()["device"] == "cpu"
        back = pickle.loads(pickle.dumps(ours))
        np.testing.assert_array_equal(back.transform(Q), got)
    with np.testing.assert_raises(ValueError):
        ours.transform(Q[:, :10])
