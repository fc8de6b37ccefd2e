"""CPU: host-side logic that needs no GPU -- result containers, centre finding, metric
resolution, argument validation, single-process mpi shim, synthetic generator."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from enspara_b200 import mpi, ra, synth
from enspara_b200.cluster import KCenters, KHybrid, KMedoids, kcenters, kmedoids, util
from enspara_b200.exception import DataInvalid, ImproperlyConfigured


def test_cluster_result_partition_np():
    """enspara/test/test_cluster_util.py:14-40."""
    r = util.ClusterResult(assignments=[0] * 20 + [1] * 20 + [2] * 20,
                           distances=[0.2] * 20 + [0.3] * 20 + [0.4] * 20,
                           center_indices=[3, 23, 43], centers=None).partition([20, 20, 20])
    assert type(r.assignments) is not ra.RaggedArray
    assert_array_equal(r.assignments[1], [1] * 20)
    assert_array_equal(r.distances[2], [0.4] * 20)
    assert_array_equal(r.center_indices, [(0, 3), (1, 3), (2, 3)])


def test_cluster_result_partition_ra():
    """enspara/test/test_cluster_util.py:43-68."""
    r = util.ClusterResult(assignments=[0] * 10 + [1] * 20 + [2] * 100,
                           distances=[0.2] * 10 + [0.3] * 20 + [0.4] * 100,
                           center_indices=[3, 23, 103], centers=None).partition([10, 20, 100])
    assert type(r.assignments) is ra.RaggedArray
    assert_array_equal(r.assignments[0], [0] * 10)
    assert_array_equal(r.assignments[2], [2] * 100)
    assert_array_equal(r.distances[1], [0.3] * 20)
    assert_array_equal(r.center_indices, [(0, 3), (1, 13), (2, 73)])
    with pytest.raises(DataInvalid):
        util.ClusterResult(assignments=[0] * 5, distances=[0.0] * 5, center_indices=[0],
                           centers=None).partition([2, 2])


def test_find_cluster_centers():
    """enspara/test/test_cluster_util.py:126-133 + first-minimum tie rule + brute force."""
    assert_array_equal(util.find_cluster_centers([1, 1, 7, 7], [.2, .1, .1, .2]), [1, 2])
    assert_array_equal(util.find_cluster_centers([0, 0, 0], [.5, .1, .1]), [1])
    with pytest.raises(DataInvalid):
        util.find_cluster_centers([0, 1], [0.1])
    rng = np.random.RandomState(0)
    a = rng.randint(0, 17, 5000)
    d = np.round(rng.rand(5000), 2)          # many exact ties
    want = []
    for c in np.unique(a):
        idx = np.where(a == c)[0]
        want.append(idx[np.argmin(d[idx])])
    assert_array_equal(util.find_cluster_centers(a, d), want)


def test_metric_resolution():
    assert util._get_distance_method("rmsd") is util.RMSD
    assert util._get_distance_method("cityblock") is util.MANHATTAN
    assert util._get_distance_method("manhattan") is util.MANHATTAN
    assert util._get_distance_method("euclidean") is util.EUCLIDEAN
    assert util._get_distance_method(util.EUCLIDEAN) is util.EUCLIDEAN
    from enspara_b200.geometry import libdist
    assert util._get_distance_method(libdist.euclidean) is util.EUCLIDEAN
    assert util._get_distance_method(libdist.manhattan) is util.MANHATTAN

    def rmsd(a, b):
        return None
    rmsd.__module__ = "mdtraj.geometry.rmsd"
    assert util._get_distance_method(rmsd) is util.RMSD
    for bad in ("nope", "chebyshev", 3, lambda X, y: X):
        with pytest.raises(ImproperlyConfigured):
            util._get_distance_method(bad)


def test_argument_validation_without_gpu():
    with pytest.raises(ImproperlyConfigured):
        KCenters(metric="rmsd")
    with pytest.raises(ImproperlyConfigured):
        KHybrid(metric="rmsd", kmedoids_updates=10)
    X = np.zeros((4, 2))
    with pytest.raises(ImproperlyConfigured):
        kcenters.kcenters(X, "euclidean")
    with pytest.raises(ImproperlyConfigured):
        kcenters.kcenters(X, "euclidean", n_clusters=None, dist_cutoff=None)
    with pytest.raises(NotImplementedError):
        kcenters.kcenters(X, "euclidean", n_clusters=2, random_first_center=True)
    with pytest.raises(ImproperlyConfigured):
        kmedoids.kmedoids(X, "euclidean")
    with pytest.raises(ImproperlyConfigured):
        kmedoids.kmedoids(X, "euclidean", cluster_center_inds=[(0, 1)])
    c = KMedoids("euclidean", n_clusters=3, n_iters=2)
    assert c.n_iters == 2 and c.metric is util.EUCLIDEAN
    est = KCenters("euclidean", n_clusters=2)
    with pytest.raises(ImproperlyConfigured):
        est.predict(X)
    assert est.mpi_mode is False


def test_single_process_mpi_shim():
    assert mpi.rank() == 0 and mpi.size() == 1
    x = np.array([3.0, 1.0, 2.0])
    assert mpi.ops.striped_array_max(x) == 3.0
    assert mpi.ops.striped_array_mean(x) == 2.0
    assert kmedoids._msq(np.array([1.0, 2.0, 3.0])) == pytest.approx(14 / 3)
    assert_array_equal(mpi.ops.distribute_frame(np.arange(6).reshape(3, 2), 1, 0), [2, 3])
    with pytest.raises(ImproperlyConfigured):
        mpi.ops.distribute_frame(np.arange(6).reshape(3, 2), 1, 1)
    rs = np.random.RandomState(3)
    ref = np.random.RandomState(3)
    for _ in range(5):
        assert mpi.ops.randind(np.arange(9), rs) == (0, ref.randint(9))
    with pytest.raises(DataInvalid):
        mpi.ops.randind(np.arange(0), rs)
    assert [int(c) for c in mpi.ops.convert_local_indices([(0, 0), (0, 7)], [3, 5])] == [0, 7]


def test_ragged_array_and_partition_helpers():
    r = ra.RaggedArray(np.arange(6), lengths=[1, 3, 2])
    assert_array_equal(r[1], [1, 2, 3])
    assert_array_equal(r[-1], [4, 5])
    assert len(r) == 3 and [len(x) for x in r] == [1, 3, 2]
    assert ra.partition_indices([0, 1, 3, 5], [1, 3, 2]) == [(0, 0), (1, 0), (1, 2), (2, 1)]
    assert [list(p) for p in ra.partition_list(list(range(6)), [1, 3, 2])] == \
        [[0], [1, 2, 3], [4, 5]]
    with pytest.raises(DataInvalid):
        ra.RaggedArray(np.arange(5), lengths=[1, 3, 2])


def test_synth_is_counter_based():
    a = synth.trajectory(300, 17, seed=4)
    b = synth.trajectory(100, 17, seed=4, first_frame=150)
    assert_array_equal(a[150:250], b)
    assert a.dtype == np.float32 and np.isfinite(a).all()
    assert not np.array_equal(a, synth.trajectory(300, 17, seed=5))
    f = synth.features(50, 8, seed=1)
    assert_array_equal(f[20:30], synth.features(10, 8, seed=1, first_row=20))
    assert f.min() >= 0 and f.max() < 1


def test_timed_logs_elapsed_seconds():
    from enspara_b200.util.log import timed
    seen = []
    with timed("took %.3f s", seen.append) as t:
        pass
    assert len(seen) == 1 and seen[0].startswith("took ") and t.seconds >= 0.0
    with pytest.raises(ValueError):            # nothing is logged when the block raises
        with timed("never %.1f", seen.append):
            raise ValueError("x")
    assert len(seen) == 1


def test_pam_draw_member_consumes_the_reference_stream():
    """PamEngine._draw_member (host only): one randint(total) per proposal like
    random_state.choice(state_inds) (kmedoids.py:514); sharded: the k-th member
    in GLOBAL frame order maps to (owner rank, k-th member on the owner), or through the
    reference's striped concatenation (mpi/ops.py:256-268) with striped_randind."""
    from types import SimpleNamespace
    from enspara_b200.cluster._pam import PamEngine
    eng = PamEngine.__new__(PamEngine)
    # single rank: cluster 0 has 7 members
    eng.shard = SimpleNamespace(size=1, rank=0)
    eng.counts_by_rank = np.array([[7, 3]])
    rs, ref = np.random.RandomState(5), np.random.RandomState(5)
    for _ in range(20):
        assert eng._draw_member(0, rs, False) == (0, int(ref.randint(7)))
    assert rs.randint(1 << 30) == ref.randint(1 << 30)        # streams still aligned
    # three ranks holding 2, 0 and 5 members of cluster 1
    eng.shard = SimpleNamespace(size=3, rank=1)
    eng.counts_by_rank = np.array([[1, 2], [1, 0], [1, 5]])
    rs, ref = np.random.RandomState(9), np.random.RandomState(9)
    seen = set()
    for _ in range(200):
        g = int(ref.randint(7))
        owner, kth = eng._draw_member(1, rs, False)
        assert (owner, kth) == ((0, g) if g < 2 else (2, g - 2))
        seen.add((owner, kth))
    assert len(seen) == 7 and all(o != 1 for o, _ in seen)     # the empty rank never owns
    # striped map: position of g in concat(arange(total)[r::size])
    rs, ref = np.random.RandomState(3), np.random.RandomState(3)
    concat = np.concatenate([np.arange(7)[r::3] for r in range(3)])
    for _ in range(50):
        g = int(np.where(concat == int(ref.randint(7)))[0][0])
        assert eng._draw_member(1, rs, True) == ((0, g) if g < 2 else (2, g - 2))
    # an empty cluster is the reference's ValueError (np.random.choice on an empty array)
    eng.counts_by_rank = np.array([[0, 2], [0, 0], [0, 5]])
    with pytest.raises(ValueError):
        eng._draw_member(0, rs, False)
