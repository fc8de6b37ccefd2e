"""GPU parity: PAM sweeps (K3 + K4 + K6), KMedoids / KHybrid, nearest-centre assignment.

Feature metrics are bit-exact against goldens made by the reference's own code.  RMSD results
must reproduce the oracle's medoid choices and assignments (near-ties below 1e-6 nm excepted)
with distances to 1e-5 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import _lib
    _lib.load()
    return torch


def test_pam_update_blobs_golden(cuda, golden):
    """enspara/test/test_cluster.py:507-530: centres [0, 7, 17], == brute-force assignment."""
    from enspara_b200.cluster import kcenters, kmedoids, util
    X = golden["blobs_X"]
    r = kcenters.kcenters(X, "sqeuclidean", n_clusters=3)
    ind, dists, assig, _ = kmedoids._kmedoids_pam_update(
        X, "sqeuclidean", r.center_indices, r.assignments, r.distances, random_state=0)
    assert_array_equal(ind, [0, 7, 17])
    assert_array_equal(assig, golden["blobs_pam_assign"])
    assert_array_equal(dists, golden["blobs_pam_dist"])
    ea, ed = util.assign_to_nearest_center(X, X[[int(i) for i in ind]], "sqeuclidean")
    assert_array_equal(assig, ea)
    assert_array_equal(dists, ed)


def test_feature_hybrid_and_kmedoids_golden(cuda, golden):
    from enspara_b200 import synth
    from enspara_b200.cluster import hybrid, kmedoids
    X = synth.features(5000, 16, seed=7)
    r = hybrid.hybrid(X, "euclidean", n_clusters=12, n_iters=3, random_state=5)
    assert [int(c) for c in r.center_indices] == golden["feat_hybrid12_centers"].tolist()
    assert_array_equal(r.assignments, golden["feat_hybrid12_assign"])
    assert_array_equal(r.distances, golden["feat_hybrid12_dist"])
    assert_array_equal(np.array(r.centers), X[golden["feat_hybrid12_centers"]])

    r = kmedoids.kmedoids(X[:600], "euclidean", n_clusters=6, n_iters=4, random_state=3)
    assert [int(c) for c in r.center_indices] == golden["feat_kmedoids6_centers"].tolist()
    assert_array_equal(r.assignments, golden["feat_kmedoids6_assign"])
    assert_array_equal(r.distances, golden["feat_kmedoids6_dist"])


def test_khybrid_zero_sweeps_equals_kcenters(cuda):
    """enspara/test/test_apps_cluster.py:509-548."""
    from sklearn.datasets import make_blobs
    from enspara_b200.cluster import KCenters, KHybrid
    X, _ = make_blobs(n_samples=100, n_features=3, random_state=3)
    a = KHybrid("euclidean", n_clusters=3, kmedoids_updates=0).fit(X)
    b = KCenters("euclidean", n_clusters=3).fit(X)
    assert_array_equal(a.distances_, b.distances_)
    assert_array_equal(a.labels_, b.labels_)


def test_rmsd_assign_matches_bruteforce(cuda):
    """enspara/test/test_cluster_util.py:88-123: assign == argmin of the full RMSD matrix."""
    from enspara_b200 import synth
    from enspara_b200.cluster import util
    from oracle import cluster as oc
    from oracle import distances as od
    for n, A, k in ((700, 22, 5), (300, 264, 21), (100, 500, 9), (50, 13, 1)):
        X = synth.trajectory(n, A, seed=A)
        T = od.Trajectory(X)
        idx = np.linspace(0, n - 1, k).astype(int)
        want_a, want_d = oc.assign_to_nearest_center(T, [T[i] for i in idx], od.rmsd)
        got_a, got_d = util.assign_to_nearest_center(T, [T[i] for i in idx], "rmsd")
        assert got_a.dtype == np.int64 and got_d.dtype == np.float64
        assert_allclose(got_d, want_d, rtol=RTOL, atol=ATOL)
        bad = np.where(got_a != want_a)[0]
        for f in bad:  # only a near-tie may differ
            dd = np.sort([od.rmsd(T[[f]], T[i])[0] for i in idx])
            assert dd[1] - dd[0] < ATOL
        assert len(bad) <= 1


def test_rmsd_pam_frame0_matches_reference_run(cuda, frame0_xyz, golden):
    """k-centers (k=3) + one PAM sweep, seed 0, on the reference's fixture; expected values
    come from the reference's own loops (scripts/make_golden.py)."""
    from enspara_b200.cluster import kcenters, kmedoids
    from oracle import distances as od
    T = od.Trajectory(frame0_xyz)
    r = kcenters.kcenters(T, "rmsd", n_clusters=3)
    ind, d, a, ctrs = kmedoids._kmedoids_pam_update(
        T, "rmsd", list(r.center_indices), r.assignments.copy(), r.distances.copy(),
        random_state=0)
    assert [int(i) for i in ind] == golden["frame0_k3_pam_centers"].tolist()
    assert_array_equal(a, golden["frame0_k3_pam_assign"])
    assert_allclose(d, golden["frame0_k3_pam_dist"], rtol=RTOL, atol=ATOL)
    assert len(ctrs) == 3 and ctrs[0].xyz.shape == (1, 22, 3)


def test_pam_update_mdtraj_golden(cuda, frame0_h5_xyz, golden):
    """The reference's own exact RMSD golden, enspara/test/test_cluster.py:533-554: k-centers
    (k=3) on frame0.h5 then one PAM sweep with seed 0 gives medoids [298, 44, 341], and the
    sweep's result equals a brute-force assignment to those medoids."""
    from enspara_b200.cluster import kcenters, kmedoids, util
    from oracle import distances as od
    X = od.Trajectory(frame0_h5_xyz)
    r = kcenters.kcenters(X, "rmsd", n_clusters=3)
    ind, dists, assig, _ = kmedoids._kmedoids_pam_update(
        X, "rmsd", r.center_indices, r.assignments, r.distances, random_state=0)
    assert_array_equal(ind, [298, 44, 341])
    expect_assig, expect_dists = util.assign_to_nearest_center(
        X, X[[int(i) for i in ind]], "rmsd")
    assert_array_equal(np.unique(assig), np.arange(3))
    assert_array_equal(assig, expect_assig)
    assert_allclose(dists, expect_dists, atol=1e-6)
    assert_array_equal(assig, golden["h5_k3_pam_assign"])
    assert_allclose(dists, golden["h5_k3_pam_dist"], rtol=RTOL, atol=ATOL)


def test_pam_update_mpi_mdtraj_golden(cuda, frame0_h5_xyz, golden):
    """enspara/test/test_cluster.py:378-419 on one rank: k=10, proposals = the first member
    of every cluster -> medoids [0, 37, 400, 105, 12, 327, 242, 346, 42, 3]."""
    from enspara_b200.cluster import kcenters, kmedoids, util
    from oracle import distances as od
    X = od.Trajectory(frame0_h5_xyz)
    r = kcenters.kcenters(X, "rmsd", n_clusters=10)
    props = [int(np.where(r.assignments == cid)[0][0]) for cid in range(10)]
    assert props == golden["h5_k10_pam_proposals"].tolist()
    ind, dists, assig, _ = kmedoids._kmedoids_pam_update(
        X, "rmsd", r.center_indices, r.assignments, r.distances, proposals=props,
        random_state=0)
    assert_array_equal(ind, [0, 37, 400, 105, 12, 327, 242, 346, 42, 3])
    true_assigs, true_dists = util.assign_to_nearest_center(
        X, X[[int(i) for i in ind]], "rmsd")
    assert_array_equal(assig, true_assigs)
    assert_allclose(dists, true_dists, rtol=1e-06, atol=1e-03)
    assert_array_equal(assig, golden["h5_k10_pam_assign"])
    assert_allclose(dists, golden["h5_k10_pam_dist"], rtol=RTOL, atol=ATOL)


def test_rmsd_hybrid_frame0(cuda, frame0_xyz, golden):
    """enspara/test/test_cluster.py:178-198 + exact comparison with the reference run."""
    from enspara_b200.cluster import hybrid
    from oracle import distances as od
    T = od.Trajectory(frame0_xyz)
    r = hybrid.hybrid(T, "rmsd", n_clusters=5, n_iters=5, random_state=0)
    assert len(np.unique(r.assignments)) == 5
    assert round(r.distances.mean(), 7) < 0.094
    assert np.std(r.distances) < 0.019
    assert [int(i) for i in r.center_indices] == golden["frame0_hybrid5_centers"].tolist()
    assert_array_equal(r.assignments, golden["frame0_hybrid5_assign"])
    assert_allclose(r.distances, golden["frame0_hybrid5_dist"], rtol=RTOL, atol=ATOL)


def test_khybrid_object_frame0(cuda, frame0_xyz):
    """enspara/test/test_cluster.py:75-112."""
    from enspara_b200.cluster import KHybrid
    from enspara_b200.exception import DataInvalid, ImproperlyConfigured
    from oracle import distances as od
    T = od.Trajectory(frame0_xyz)
    with pytest.raises(ImproperlyConfigured):
        KHybrid(metric="rmsd", kmedoids_updates=10)
    c = KHybrid(metric="rmsd", n_clusters=5, kmedoids_updates=10, random_state=99).fit(T)
    assert len(np.unique(c.labels_)) == 5
    assert abs(np.average(c.distances_) - 0.08) < 0.005
    assert abs(np.std(c.distances_) - 0.0185) < 0.005
    with pytest.raises(DataInvalid):
        c.result_.partition([5, 10])
    p = c.result_.partition([len(T) - 100, 100])
    assert np.all(p.distances[0] == c.distances_[0:-100])
    assert np.all(p.assignments[0] == c.labels_[0:-100])
    assert len(p.center_indices[0]) == 2


def test_rmsd_hybrid_synthetic_matches_oracle(cuda):
    """BASELINE config 3 shape (500 atoms) scaled to what the oracle does in seconds."""
    from enspara_b200 import synth
    from enspara_b200.cluster import hybrid
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.trajectory(1500, 500, seed=4)
    T = od.Trajectory(X)
    ref = oc.hybrid(T, od.rmsd, n_clusters=12, n_iters=2, random_state=0)
    got = hybrid.hybrid(T, "rmsd", n_clusters=12, n_iters=2, random_state=0)
    assert [int(i) for i in got.center_indices] == [int(i) for i in ref.center_indices]
    assert_array_equal(got.assignments, ref.assignments)
    assert_allclose(got.distances, ref.distances, rtol=RTOL, atol=ATOL)


def test_kmedoids_warm_start_and_proposals(cuda, frame0_xyz):
    """enspara/test/test_cluster.py:137-176."""
    import copy
    from enspara_b200.cluster import kcenters, kmedoids
    from oracle import distances as od
    T = od.Trajectory(frame0_xyz)
    proposals = np.random.RandomState(1).randint(0, len(T), 5)
    r = kcenters.kcenters(T, "rmsd", n_clusters=5)
    ci, a, d = copy.deepcopy(r.center_indices), r.assignments.copy(), r.distances.copy()
    ci2, d2, a2, _ = kmedoids._kmedoids_pam_update(
        T, "rmsd", r.center_indices, r.assignments, r.distances, proposals=proposals,
        random_state=0)
    rr = kmedoids.kmedoids(T, "rmsd", 5, n_iters=1, assignments=a, distances=d,
                           cluster_center_inds=ci, proposals=proposals, random_state=0)
    assert_allclose(d2, rr.distances, atol=1e-3)
    assert_array_equal(a2, rr.assignments)
    assert_array_equal(ci2, rr.center_indices)


def test_kmedoids_input_errors(cuda):
    from enspara_b200.cluster import kmedoids
    from enspara_b200.exception import DataInvalid, ImproperlyConfigured
    X = np.random.RandomState(0).rand(30, 2)
    with pytest.raises(ImproperlyConfigured):
        kmedoids.kmedoids(X, "euclidean")
    with pytest.raises(ImproperlyConfigured):
        kmedoids.kmedoids(X, "euclidean", n_clusters=2, assignments=np.zeros(30, int))
    with pytest.raises(ImproperlyConfigured):
        kmedoids.kmedoids(X, "euclidean", cluster_center_inds=[(0, 1)])
    with pytest.raises(DataInvalid):
        kmedoids._kmedoids_pam_update(X, "euclidean", [0, 1], np.zeros(30, int),
                                      np.zeros(30), proposals=[1, 2, 3])


def test_pam_pruned_full_pass_changes_nothing(cuda):
    """The triangle-inequality pruning of the PAM full pass (frames that provably stay with
    their medoid are not read) gives bit-identical sweeps, and so does queueing a whole
    proposal with one C call instead of step by step."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters, util
    from enspara_b200.cluster._pam import PamEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    from enspara_b200.device import DeviceTrajectory
    all_forms = ((True, True, True, "list"), (True, True, True, "all"),
                 (True, True, True, "cap1"), (True, True, False, None),
                 (True, False, False, None), (False, False, False, None))
    # the last case has more medoids than one 1024-wide chunk of the list kernel, and more than
    # the list's capacity
    for n, A, k, forms in ((6000, 40, 50, all_forms), (3000, 22, 100, all_forms),
                           (501, 22, 7, all_forms), (12000, 22, 1300, all_forms[:2])):
        data = DeviceTrajectory.from_host(synth.trajectory(n, A, seed=n))
        r = kcenters.kcenters(data, "rmsd", n_clusters=k)
        ctr = [int(c) for c in r.center_indices]
        out = []
        for prune, compact, one_call, med_list in forms:
            # default: the whole proposal queued by ONE C call (eb_pam_propose_rmsd) with the
            # ambiguous frames re-assigned against the medoids the triangle inequality leaves;
            # the same call against all medoids; a medoid list of capacity 1 (overflows ->
            # general path, then the engine stops listing); then the steps issued one by one
            # from Python, the fused skip-rounds pruned pass, and the unpruned full pass
            pam = PamEngine(data, util.RMSD, _SingleComm(), r.distances, r.assignments, ctr)
            assert pam.prune and pam.prune_compact and pam._ctx is not None
            assert pam._ctx.use_list == 1
            pam.prune = prune
            pam.prune_compact = compact
            if med_list == "all":
                pam._ctx.use_list = 0
            elif med_list == "cap1":
                pam._ctx.med_list_cap = 1
            if not one_call:
                pam._ctx = None
            log = []
            for sweep in range(2 if k < 1000 else 1):
                pam.sweep(random_state=sweep, log=log)
            a, d = pam.results_host()
            out.append((list(pam.medoid_global), a, d, log))
        for other in out[1:]:
            assert out[0][0] == other[0]
            assert_array_equal(out[0][1], other[1])
            assert_array_equal(out[0][2], other[2])
            assert out[0][3] == other[3]     # same proposals, same costs, same decisions


def test_kmedoids_update_mpi_numpy_golden(cuda):
    """enspara/test/test_cluster.py:422-463 on one rank: blobs (random_state=1, 20 samples),
    squared euclid, proposals = the first point of every true label -> medoids [0, 3, 19]."""
    from sklearn.datasets import make_blobs
    from enspara_b200.cluster import kcenters, kmedoids, util
    X, y = make_blobs(centers=[(0, 0), (0, 10), (10, 0)], random_state=1, n_samples=20)
    r = kcenters.kcenters(X, "sqeuclidean", n_clusters=3)
    props = [int(np.where(y == cid)[0][0]) for cid in range(3)]
    ind, dists, assig, _ = kmedoids._kmedoids_pam_update(
        X, "sqeuclidean", r.center_indices, r.assignments, r.distances, proposals=props,
        random_state=0)
    assert_array_equal(ind, [0, 3, 19])
    ea, ed = util.assign_to_nearest_center(X, X[[int(i) for i in ind]], "sqeuclidean")
    assert_array_equal(assig, ea)
    assert_allclose(dists, ed, rtol=1e-06, atol=1e-03)
