"""GPU: edge cases of the clustering path -- degenerate sizes, more clusters requested than
frames, duplicate frames (exact ties), one centre, more centres than frames (the reference's
alternate branch, cluster/util.py:193-197)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def test_more_clusters_than_frames(cuda):
    """n_clusters > n: the loop stops when every frame is a centre (maxdist reaches 0,
    kcenters.py:217)."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.trajectory(7, 12, seed=1)
    T = od.Trajectory(X)
    ref = oc.kcenters(T, od.rmsd, n_clusters=50)
    got = kcenters.kcenters(T, "rmsd", n_clusters=50)
    # once every frame is a centre the remaining "distances" are self-distances: rounding noise
    # (exactly 0 here, ~1e-8 in the float64 oracle, ~1e-4 in mdtraj's float32), i.e. the
    # documented sub-1e-6 nm regime -- the first n picks are what is defined
    gc, rc = [int(c) for c in got.center_indices], [int(c) for c in ref.center_indices]
    assert gc[:7] == rc[:7] and sorted(gc[:7]) == list(range(7))
    assert got.distances.max() < 1e-6
    F = synth.features(9, 5, seed=2)
    ref = oc.kcenters(F, od.euclidean, n_clusters=50)
    got = kcenters.kcenters(F, "euclidean", n_clusters=50)
    assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices]
    assert len(got.center_indices) == 9
    assert_array_equal(got.distances, ref.distances)


def test_duplicate_frames_exact_ties(cuda):
    """Duplicated rows give exact distance ties: the first occurrence wins the arg-max
    (np.argmax, kcenters.py:282) and the earlier centre keeps a tied frame (strict '<', :304)."""
    from enspara_b200.cluster import kcenters, util
    from oracle import cluster as oc
    from oracle import distances as od
    rs = np.random.RandomState(0)
    base = rs.rand(40, 6).astype(np.float32)
    F = np.concatenate([base, base, base[:10]])          # every row appears 2-3 times
    for metric, om in (("euclidean", od.euclidean), ("manhattan", od.manhattan)):
        ref = oc.kcenters(F, om, n_clusters=25)
        got = kcenters.kcenters(F, metric, n_clusters=25)
        assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices]
        assert_array_equal(got.assignments, ref.assignments)
        assert_array_equal(got.distances, ref.distances)
    # nearest-centre assignment with duplicated centres: lowest centre index wins
    centers = F[[3, 43, 7, 3]]
    a, d = util.assign_to_nearest_center(F, centers, "euclidean")
    ea, ed = oc.assign_to_nearest_center(F, centers, od.euclidean)
    assert_array_equal(a, ea)
    assert_array_equal(d, ed)
    assert a[3] == 0 and a[43] == 0 and a[83] == 0
    # the same with coordinates: duplicated frames, RMSD
    from enspara_b200 import synth
    X = synth.trajectory(30, 20, seed=3)
    X = np.concatenate([X, X[:15]])
    T = od.Trajectory(X)
    ref = oc.kcenters(T, od.rmsd, n_clusters=12)
    got = kcenters.kcenters(T, "rmsd", n_clusters=12)
    assert [int(c) for c in got.center_indices] == [int(c) for c in ref.center_indices]
    assert_array_equal(got.assignments, ref.assignments)


def test_single_frame_and_single_center(cuda):
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters, util
    X = synth.trajectory(1, 30, seed=4)
    r = kcenters.kcenters(X, "rmsd", n_clusters=3)
    assert [int(c) for c in r.center_indices] == [0]
    assert_array_equal(r.assignments, [0])
    assert r.distances[0] < 1e-4
    Y = synth.trajectory(200, 30, seed=5)
    a, d = util.assign_to_nearest_center(Y, Y[[17]], "rmsd")
    assert_array_equal(a, np.zeros(200, dtype=np.int64))
    assert_allclose(d, util.RMSD(Y, Y[17]), rtol=0, atol=0)


def test_more_centers_than_frames(cuda):
    """cluster/util.py:193-197 loops over frames with argmin (first minimum) when there are
    more centres than frames; the result equals the main branch's lowest-index rule."""
    from enspara_b200 import synth
    from enspara_b200.cluster import util
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.trajectory(300, 25, seed=6)
    frames = od.Trajectory(X[:20])
    centers = od.Trajectory(X[10:300])            # 290 centres > 20 frames, 10 of them frames
    ea, ed = oc.assign_to_nearest_center(frames, centers, od.rmsd)
    a, d = util.assign_to_nearest_center(frames, centers, "rmsd")
    assert_array_equal(a, ea)
    assert_allclose(d, ed, rtol=1e-5, atol=1e-6)
    assert_array_equal(a[10:], np.arange(10))


def test_empty_subset_and_zero_iterations(cuda):
    from enspara_b200 import synth
    from enspara_b200.cluster import KHybrid, kcenters
    X = synth.trajectory(400, 16, seed=7)
    r0 = kcenters.kcenters(X, "rmsd", n_clusters=5)
    h = KHybrid("rmsd", n_clusters=5, kmedoids_updates=0, random_state=0).fit(X)
    assert [int(c) for c in h.result_.center_indices] == [int(c) for c in r0.center_indices]
    assert_array_equal(h.result_.distances, r0.distances)
