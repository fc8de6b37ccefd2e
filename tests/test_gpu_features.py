"""GPU parity: libdist replacements (K2) and feature-vector clustering -- BIT-EXACT bar.

The reference arithmetic (typed difference, typed square, float64 accumulation in feature order,
double sqrt: enspara/geometry/libdist.pyx:100-145) is reproduced exactly, so every comparison
here is assert_array_equal against the oracle and against goldens produced by the reference's
own code (tests/golden/reference_runs.npz, scripts/make_golden.py).
"""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import _lib
    _lib.load()
    return torch


def test_libdist_reference_cases(cuda):
    """enspara/test/test_libdist.py:34-108 (== scipy cdist exactly, error cases)."""
    from scipy.spatial.distance import cdist
    from enspara_b200.exception import DataInvalid
    from enspara_b200.geometry import libdist
    X = np.array([[1, 1], [2, 2], [3, 3], [-1, 3]])
    y = np.array([0, 0])
    for fn, name in ((libdist.euclidean, "euclidean"), (libdist.manhattan, "cityblock")):
        with pytest.raises(DataInvalid):
            fn(X, y.reshape(1, -1))
        with pytest.raises(DataInvalid):
            fn(X.reshape(1, -1), y)
        with pytest.raises(DataInvalid):
            fn(X.flatten(), y)
        with pytest.raises(DataInvalid):
            fn(X, y[1:])
        assert_array_equal(fn(X, y), cdist(X, y.reshape(1, -1), metric=name).flatten())
    with pytest.raises(DataInvalid):
        libdist.euclidean(X, y, out=np.empty(shape=(X.shape[0]), dtype="int"))
    with pytest.raises(DataInvalid):
        libdist.euclidean(X, y, out=np.empty(shape=(X.shape[0] - 1)))
    d = libdist.euclidean(X, y, out=np.empty(shape=(X.shape[0]), dtype="float64"))
    assert_array_equal(d, cdist(X, y.reshape(1, -1)).flatten())
    with pytest.raises(ValueError):  # mixed dtypes
        libdist.euclidean(X.astype(np.float32), y.astype(np.float64))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int8, np.int16, np.int32, np.int64])
@pytest.mark.parametrize("F", [1, 3, 64, 65, 200])
def test_one_to_all_bit_exact(cuda, dtype, F):
    from enspara_b200.geometry import libdist
    from oracle import distances as od
    rng = np.random.default_rng(F)
    n = 1000 + F
    if np.issubdtype(dtype, np.integer):
        X = rng.integers(-100, 100, (n, F)).astype(dtype)
    else:
        X = (rng.random((n, F)) * 10 - 5).astype(dtype)
    y = X[n // 3].copy()
    assert_array_equal(libdist.euclidean(X, y), od.euclidean(X, y))
    assert_array_equal(libdist.manhattan(X, y), od.manhattan(X, y))


def test_kcenters_euclidean_golden(cuda, golden):
    """Reference run end to end (its Python loop + its Cython libdist) -- exact."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters
    X = synth.features(5000, 16, seed=7)
    r = kcenters.kcenters(X, "euclidean", n_clusters=40)
    assert [int(c) for c in r.center_indices] == golden["feat_k40_centers"].tolist()
    assert_array_equal(r.assignments, golden["feat_k40_assign"])
    assert_array_equal(r.distances, golden["feat_k40_dist"])
    assert r.distances.dtype == np.float64 and r.assignments.dtype == np.int64
    assert_array_equal(np.array(r.centers), X[golden["feat_k40_centers"]])

    r = kcenters.kcenters(X, "euclidean", dist_cutoff=0.9)
    assert [int(c) for c in r.center_indices] == golden["feat_cut09_centers"].tolist()
    assert_array_equal(r.distances, golden["feat_cut09_dist"])

    r = kcenters.kcenters(X, "manhattan", n_clusters=25)
    assert [int(c) for c in r.center_indices] == golden["feat_manh_k25_centers"].tolist()
    assert_array_equal(r.assignments, golden["feat_manh_k25_assign"])
    assert_array_equal(r.distances, golden["feat_manh_k25_dist"])


def test_assign_to_nearest_center_golden(cuda, golden):
    from enspara_b200 import synth
    from enspara_b200.cluster import util
    from enspara_b200.geometry import libdist
    X = synth.features(5000, 16, seed=7)
    a, d = util.assign_to_nearest_center(X, X[[5, 17, 99, 1234, 4000]], libdist.euclidean)
    assert_array_equal(a, golden["feat_assign5_assign"])
    assert_array_equal(d, golden["feat_assign5_dist"])


def test_kcenters_c2_shape_matches_oracle(cuda):
    """BASELINE config 2 shape (64-dim float32), 50k rows, exact against the oracle."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.features(50000, 64, seed=2)
    ref = oc.kcenters(X, od.euclidean, n_clusters=30)
    got = kcenters.kcenters(X, "euclidean", n_clusters=30)
    assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices]
    assert_array_equal(got.assignments, ref.assignments)
    assert_array_equal(got.distances, ref.distances)


def test_kcenters_hot_start_and_predict(cuda):
    """enspara/test/test_cluster.py:593-637."""
    from enspara_b200.cluster import KCenters
    from enspara_b200.cluster.util import ClusterResult
    from oracle import cluster as oc
    from oracle import distances as od
    rg = np.random.RandomState(seed=41)
    gens = [(1, 1), (10, 10), (0, 20)]
    X = np.concatenate([np.stack([rg.normal(g[0], 1, 20), rg.normal(g[1], 1, 20)], axis=1)
                        for g in gens])
    clust = KCenters(metric="euclidean", cluster_radius=6)
    init = np.array(gens[0:2], dtype=float)
    clust.fit(X=X, init_centers=init)
    ref = oc.kcenters(X, od.euclidean, dist_cutoff=6, init_centers=init)
    assert [int(c) for c in clust.result_.center_indices] == [int(c) for c in ref.center_indices]
    assert_array_equal(clust.result_.assignments, ref.assignments)
    assert_array_equal(clust.result_.distances, ref.distances)
    assert len(clust.result_.center_indices) == 3

    centers = np.array(gens, dtype="float64")
    clust = KCenters(metric="euclidean", cluster_radius=2)
    clust.result_ = ClusterResult(centers=centers, assignments=None, distances=None,
                                  center_indices=None)
    p = clust.predict(X)
    assert_array_equal(p.assignments, [0] * 20 + [1] * 20 + [2] * 20)
    assert np.all(p.distances < 4)
    assert np.argmin(p.distances[0:20]) == p.center_indices[0]
    assert p.centers is centers


def test_unknown_callable_rejected(cuda):
    from enspara_b200.cluster import kcenters
    from enspara_b200.exception import ImproperlyConfigured
    with pytest.raises(ImproperlyConfigured):
        kcenters.kcenters(np.zeros((4, 2)), lambda X, y: X[:, 0], n_clusters=2)
    with pytest.raises(ImproperlyConfigured):
        kcenters.kcenters(np.zeros((4, 2)), "not-a-metric", n_clusters=2)
    with pytest.raises(ImproperlyConfigured):
        kcenters.kcenters(np.zeros((4, 2)), "euclidean")
    with pytest.raises(NotImplementedError):
        kcenters.kcenters(np.zeros((4, 2)), "euclidean", n_clusters=2, random_first_center=True)


# ---------------------------------------------------------------------------------------------
# persistent multi-iteration step (csrc/eb_feat.cu k_kcenters_multi_feat): rows of whole
# 128-byte boxes on a single shard run ALL iterations of a batch in one cooperative launch
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,F", [(np.float32, 64), (np.float32, 32), (np.float32, 96),
                                     (np.float64, 16), (np.float64, 48), (np.int8, 128),
                                     (np.int16, 64), (np.int32, 32), (np.int64, 16)])
def test_multi_iteration_kernel_bit_exact(cuda, dtype, F):
    """n_clusters-bounded, cutoff-terminated (stop inside a 32-launch batch and exactly at a
    batch boundary) and both bounds at once -- centres, assignments, float64 distances equal to
    the oracle's (= the reference's own compiled libdist, tests/test_oracle_ref.py)."""
    from enspara_b200.cluster import kcenters
    from oracle import cluster as oc
    from oracle import distances as od
    rng = np.random.default_rng(int(F) + np.dtype(dtype).itemsize)
    n = 20_011
    if np.issubdtype(dtype, np.integer):
        X = rng.integers(-100, 100, (n, F)).astype(dtype)
    else:
        X = rng.random((n, F)).astype(dtype)
    trace = []
    ref = oc.kcenters(X, od.euclidean, n_clusters=70, trace=trace)
    got = kcenters.kcenters(X, "euclidean", n_clusters=70)
    assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices]
    assert_array_equal(got.assignments, ref.assignments)
    assert_array_equal(got.distances, ref.distances)
    # cutoff between the 40th and 41st max-min-distance (inside the second batch of 32), at
    # the 32nd (batch boundary) and at the 1st
    for stop_after in (40, 32, 1):
        cut = 0.5 * (trace[stop_after - 1][1] + trace[stop_after][1]) \
            if trace[stop_after - 1][1] > trace[stop_after][1] else trace[stop_after - 1][1]
        ref_c = oc.kcenters(X, od.euclidean, dist_cutoff=cut)
        got_c = kcenters.kcenters(X, "euclidean", dist_cutoff=cut)
        assert [int(c) for c in got_c.center_indices] == [int(c) for c in ref_c.center_indices]
        assert_array_equal(got_c.assignments, ref_c.assignments)
        assert_array_equal(got_c.distances, ref_c.distances)
    ref_b = oc.kcenters(X, od.manhattan, n_clusters=45, dist_cutoff=float(trace[50][1]))
    got_b = kcenters.kcenters(X, "manhattan", n_clusters=45, dist_cutoff=float(trace[50][1]))
    assert [int(c) for c in got_b.center_indices] == [int(c) for c in ref_b.center_indices]
    assert_array_equal(got_b.distances, ref_b.distances)


def test_multi_iteration_kernel_edge_cases(cuda):
    """More clusters requested than rows (every row becomes a centre, then the stop rule fires
    on maxdist == 0), duplicate rows (exact ties -> first occurrence), a warm start
    (init_centers: centre ids continue after the existing ones), and a tiny shard whose grid is
    a single block."""
    from enspara_b200.cluster import kcenters
    from oracle import cluster as oc
    from oracle import distances as od
    rs = np.random.RandomState(5)
    base = rs.rand(300, 64).astype(np.float32)
    for X, kw in ((base[:9], dict(n_clusters=50)),
                  (np.concatenate([base, base, base[:50]]), dict(n_clusters=120)),
                  (base, dict(n_clusters=40, init_centers=base[[7, 100, 250]])),
                  (base[:33], dict(dist_cutoff=2.5))):
        ref = oc.kcenters(X, od.euclidean, **kw)
        got = kcenters.kcenters(X, "euclidean", **kw)
        assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices], kw
        assert_array_equal(got.assignments, ref.assignments)
        assert_array_equal(got.distances, ref.distances)


def test_multi_and_single_launches_mix(cuda):
    """The engine queues batches; a batch of one step takes the single-launch kernel, longer
    ones the persistent kernel.  Driving the engine with batch sizes 1, 5, 1, 32, 3 must give
    the run a plain call gives."""
    import torch
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters as kc
    from enspara_b200.cluster._engine import KCentersEngine
    X = synth.device_features(30_000, 64, seed=4)
    want = kc.kcenters(X, "euclidean", n_clusters=42)
    eng = KCentersEngine(X, "euclidean", kc._SingleComm())
    eng._ensure_center_list(64)
    eng.seed(0)
    for b in (1, 5, 1, 32, 3):
        eng.step(1 << 30, 0.0, b)
    st = eng.read_state()
    assert st.n_centers == 42 and not st.done
    got_c = eng.center_list[:42].cpu().numpy().tolist()
    assert got_c == [int(c) for c in want.center_indices]
    a, d = eng.results_host()
    assert_array_equal(a, want.assignments)
    assert_array_equal(d, want.distances)
    # a limit reached inside a batch: the remaining launches are no-ops
    eng2 = KCentersEngine(X, "euclidean", kc._SingleComm())
    eng2._ensure_center_list(64)
    eng2.seed(0)
    eng2.step(10, 0.0, 25)
    st2 = eng2.read_state()
    assert st2.n_centers == 10 and st2.done == 1
    eng2.step(10, 0.0, 4)
    st3 = eng2.read_state()
    assert st3.n_centers == 10 and st3.done == 1 and st3.n_noop >= st2.n_noop + 4
    assert eng2.center_list[:10].cpu().numpy().tolist() == got_c[:10]
    torch.cuda.synchronize()
