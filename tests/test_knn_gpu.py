"""GPU parity tests of the nearest-neighbour engine against the golden vectors of the unmodified
NearestNeighborsLinear and against the oracle: identical index sets except distance ties, distances within 1e-6 rel
(the reference's own test compares distances with eps 1e-6, tests/datastructures/nearestneighbors.cpp:69,149-158)."""
import os

import numpy as np
import pytest

import ompl_b200
from oracle import port

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "knn_linear.npz")


def assert_knn_parity(idx, dist, ridx, rdist):
    fin = np.isfinite(rdist)
    assert (np.isfinite(dist) == fin).all()
    np.testing.assert_allclose(dist[fin], rdist[fin], rtol=1e-6, atol=0)
    # identical neighbour sets per query except members tied at the boundary distance
    n_tie = 0
    for i in range(idx.shape[0]):
        a, b = set(idx[i][idx[i] >= 0]), set(ridx[i][ridx[i] >= 0])
        if a != b:
            kth = rdist[i][fin[i]].max()
            for j in a ^ b:
                dj = np.concatenate([dist[i][idx[i] == j], rdist[i][ridx[i] == j]])[0]
                assert abs(dj - kth) <= 1e-6 * max(kth, 1e-300), (i, j, dj, kth)
            n_tie += 1
    return n_tie


@pytest.mark.parametrize("tag", ["d7", "d8", "d20", "d3"])
def test_golden_linear(ctx, tag):
    z = np.load(GOLDEN)
    db, qs, k = z[tag + "_db"], z[tag + "_q"], int(z[tag + "_k"])
    nn = ompl_b200.NearestNeighbors(ctx, db.shape[1])
    nn.add(db)
    assert nn.size() == len(db)
    idx, dist, cnt = nn.nearest_k(qs, k)
    np.testing.assert_array_equal(cnt, z[tag + "_cnt"])          # k > n returns all n
    assert_knn_parity(idx, dist, z[tag + "_idx"], z[tag + "_dist"])
    np.testing.assert_array_equal(dist, z[tag + "_dist"])        # exact fp64 arithmetic of the reference metric
    rc = [len(nn.nearest_r(q, 2.0)[0]) for q in qs]
    np.testing.assert_array_equal(rc, z[tag + "_rcnt"])
    off = 0
    for q, c in zip(qs, z[tag + "_rcnt"]):
        ii, dd = nn.nearest_r(q, 2.0)
        assert (np.diff(dd) >= 0).all()
        assert set(ii) == set(z[tag + "_ridx"][off:off + c])
        off += c


@pytest.mark.parametrize("dim,n,nq,k", [(7, 20000, 3000, 16), (8, 5000, 700, 10), (14, 4000, 300, 32), (7, 700, 5, 1), (2, 3000, 257, 5)])
def test_seeded_vs_oracle(ctx, dim, n, nq, k):
    rng = np.random.default_rng(dim * 1000 + n)
    db = rng.uniform(-3, 3, (n, dim))
    qs = rng.uniform(-3, 3, (nq, dim))
    nn = ompl_b200.NearestNeighbors(ctx, dim)
    nn.add(db[: n // 2])
    nn.add(db[n // 2:])          # incremental add
    idx, dist, cnt = nn.nearest_k(qs, k)
    ridx, rdist = port.knn(db, qs, k, nthreads=8)
    assert (cnt == k).all()
    assert_knn_parity(idx, dist, ridx, rdist)
    np.testing.assert_array_equal(idx, ridx)
    assert (np.diff(dist, axis=1) >= 0).all()


def test_float32_origin_points_and_exclude_self(ctx):
    # roadmap case: database = fp32 samples, every point queries the set it belongs to, itself excluded (PRM.cpp:583,601)
    rng = np.random.default_rng(3)
    db = rng.uniform(-3, 3, (6000, 7)).astype(np.float32).astype(np.float64)
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.add(db)
    idx, dist, _ = nn.nearest_k(db, 16, self_ids=np.arange(len(db)))
    ridx, rdist = port.knn(db, db, 16, exclude_self=True, nthreads=8)
    assert_knn_parity(idx, dist, ridx, rdist)
    assert not (idx == np.arange(len(db))[:, None]).any()
    # without exclusion a member query returns itself first at distance 0 (nearestneighbors.cpp:172-174)
    idx2, dist2, _ = nn.nearest_k(db[:100], 4)
    np.testing.assert_array_equal(idx2[:, 0], np.arange(100))
    assert (dist2[:, 0] == 0).all()


def test_ties_and_duplicates(ctx):
    rng = np.random.default_rng(9)
    base = rng.integers(-2, 3, (400, 3)).astype(np.float64)   # lattice points: many exact ties and duplicates
    nn = ompl_b200.NearestNeighbors(ctx, 3)
    nn.add(base)
    qs = rng.integers(-2, 3, (50, 3)).astype(np.float64)
    idx, dist, _ = nn.nearest_k(qs, 12)
    ridx, rdist = port.knn(base, qs, 12)
    np.testing.assert_array_equal(dist, rdist)        # distances identical even where identities are tied
    np.testing.assert_array_equal(idx, ridx)          # and ties are resolved by id on both sides


def test_remove_clear_and_empty(ctx):
    rng = np.random.default_rng(4)
    db = rng.uniform(-1, 1, (300, 7))
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.add(db)
    q = db[17] + 1e-3
    assert nn.nearest(q) == 17
    nn.remove(17)
    assert nn.size() == 299
    idx, dist, _ = nn.nearest_k(q, 299 + 5)
    assert 17 not in idx[0] and (idx[0] >= 0).sum() == 299
    keep = np.delete(np.arange(300), 17)
    ridx, rdist = port.knn(db[keep], q[None], 299)
    np.testing.assert_array_equal(keep[ridx[0]], idx[0][:299])
    ii, dd = nn.nearest_r(q, np.inf)                  # nearestR(inf) returns everything, sorted (nearestneighbors.cpp:160-170)
    assert len(ii) == 299 and 17 not in ii and (np.diff(dd) >= 0).all()
    with pytest.raises(ompl_b200.B200MPError):
        nn.remove(17)
    nn.clear()
    assert nn.size() == 0
    with pytest.raises(ompl_b200.B200MPError) as e:   # "No elements found ..." (NearestNeighborsLinear.h:115)
        nn.nearest(q)
    assert "No elements found" in str(e.value)
    idx, dist, cnt = nn.nearest_k(q, 3)
    assert cnt[0] == 0 and (idx == -1).all()


def test_single_query_large_database(ctx):
    rng = np.random.default_rng(5)
    db = rng.uniform(-3, 3, (200000, 7))
    nn = ompl_b200.NearestNeighbors(ctx, 7)
    nn.add(db)
    q = rng.uniform(-3, 3, (1, 7))
    idx, dist, _ = nn.nearest_k(q, 16)
    ridx, rdist = port.knn(db, q, 16, nthreads=8)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_allclose(dist, rdist, rtol=1e-12)  # (device sincos vs libm: last-bit differences in the link tables)


@pytest.mark.parametrize("dim,n,nq,k,kind", [(7, 40000, 40000, 16, "self"), (8, 50000, 6000, 10, "uniform"), (14, 33000, 5000, 32, "uniform"),
                                              (7, 70000, 9000, 16, "clustered"), (3, 40000, 5000, 8, "duplicates")])
def test_cell_index_equals_exhaustive_search(ctx, dim, n, nq, k, kind):
    """Large batches go through the cell index (pruned, exact); the exhaustive tiled kernel is the same engine with the
    index switched off, itself pinned on the oracle above. Results must be identical, ids and fp64 distances."""
    rng = np.random.default_rng(dim * 7 + n)
    if kind == "clustered":
        centres = rng.uniform(-3, 3, (20, dim))
        db = centres[rng.integers(0, 20, n)] + rng.normal(scale=0.05, size=(n, dim))
    elif kind == "duplicates":
        db = rng.integers(-3, 4, (n, dim)).astype(np.float64)      # 343 distinct points, thousands of exact ties
    else:
        db = rng.uniform(-3, 3, (n, dim))
    db = db.astype(np.float32).astype(np.float64) if kind == "self" else db
    qs = db[:nq] if kind == "self" else (db[rng.integers(0, n, nq)] + rng.normal(scale=0.3, size=(nq, dim)))
    self_ids = np.arange(nq) if kind == "self" else None
    nn = ompl_b200.NearestNeighbors(ctx, dim)
    nn.add(db)
    idx, dist, cnt = nn.nearest_k(qs, k, self_ids=self_ids)
    nn.set_index(False)
    ridx, rdist, rcnt = nn.nearest_k(qs, k, self_ids=self_ids)
    np.testing.assert_array_equal(cnt, rcnt)
    np.testing.assert_array_equal(dist, rdist)
    np.testing.assert_array_equal(idx, ridx)
    # and a sample against the CPU oracle
    sub = rng.choice(nq, 64, replace=False)
    oidx, odist = port.knn(db, qs[sub], k + 1, nthreads=8)
    for r, i in enumerate(sub):
        keep = np.ones(k + 1, bool)
        if self_ids is not None:
            keep = oidx[r] != self_ids[i]
        np.testing.assert_array_equal(dist[i], odist[r][keep][:k])
    # removal invalidates the index; the rebuilt one skips the tombstone
    nn.set_index(True)
    victim = int(idx[0, 0])
    nn.remove(victim)
    idx2, _, _ = nn.nearest_k(qs[:5000], k, self_ids=None if self_ids is None else self_ids[:5000])
    assert victim not in idx2


@pytest.mark.parametrize("n,dim,k", [(100000, 7, 6000), (5000, 3, 4999), (3000, 6, 5000)])
def test_large_k_is_selected_on_the_device(ctx, n, dim, k):
    """k beyond the kernels' list sizes (RRT* / PRM* style k ~ log n, or "everything, sorted"): radix select + sort on the device.
    Same answer as a full sort: ids ascending by (distance, id), distances the reference's doubles, missing entries -1 / +inf."""
    rng = np.random.default_rng(n + k)
    db = rng.uniform(-2, 2, (n, dim))
    db[7] = db[3]                       # an exact duplicate: ties are ordered by id
    nn = ompl_b200.NearestNeighbors(ctx, dim)
    nn.add(db)
    nn.remove(11)
    q = rng.uniform(-2, 2, (2, dim))
    q[1] = db[5]
    before = ctx.launch_count
    idx, dist, cnt = nn.nearest_k(q, k, self_ids=np.array([-1, 5]))
    assert ctx.launch_count - before > 20   # the device path, not a host sort
    for i in range(2):
        d2 = np.zeros(n)
        for j in range(dim):
            diff = db[:, j] - q[i, j]
            d2 += diff * diff
        d = np.sqrt(d2)
        d[11] = np.inf
        if i == 1:
            d[5] = np.inf
        order = np.lexsort((np.arange(n), d))
        live = int(np.isfinite(d).sum())
        m = min(k, live)
        assert cnt[i] == m
        np.testing.assert_array_equal(idx[i, :m], order[:m])
        np.testing.assert_array_equal(dist[i, :m], d[order[:m]])
        assert (idx[i, m:] == -1).all() and np.isinf(dist[i, m:]).all()


def test_large_k_chain_metric(ctx):
    n, links = 4000, 20
    rng = np.random.default_rng(1)
    db = rng.uniform(-np.pi, np.pi, (n, links))
    nn = ompl_b200.NearestNeighbors(ctx, links, metric=1, link_length=1.0 / links)
    nn.add(db)
    idx, dist, cnt = nn.nearest_k(db[:1], 1000)
    ridx, rdist = port.knn(db, db[:1], 1000, metric=1, param=1.0 / links, nthreads=4)
    assert cnt[0] == 1000
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_allclose(dist, rdist, rtol=1e-12)  # (device sincos vs libm: last-bit differences in the link tables)
