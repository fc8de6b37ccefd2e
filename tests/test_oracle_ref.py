"""CPU, build container only: the numpy/C restatement in oracle/ against the REAL reference
package imported from /root/reference (its own Python loops, its own Cython libdist).
Skipped where the reference tree is absent (e.g. on the GPU box)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle import cluster as oc
from oracle import distances as od
from oracle import refharness as rh

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    import logging
    logging.disable(logging.CRITICAL)
    mods = rh.modules()
    yield mods
    logging.disable(logging.NOTSET)


def test_libdist_restatement_bit_exact(ref):
    libdist = ref[4]
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        X = (rng.random((2000, 64)) * 4 - 2).astype(dt)
        y = X[3].copy()
        assert_array_equal(libdist.euclidean(X, y), od.euclidean(X, y))
        assert_array_equal(libdist.manhattan(X, y), od.manhattan(X, y))
    for dt in (np.int8, np.int16, np.int32, np.int64):
        X = rng.integers(-50, 50, (100, 7)).astype(dt)
        y = X[5].copy()
        assert_array_equal(libdist.euclidean(X, y), od.euclidean(X, y))
        assert_array_equal(libdist.manhattan(X, y), od.manhattan(X, y))


def test_kcenters_restatement(ref, frame0_xyz):
    kc = ref[0]
    from enspara_b200 import synth
    X = synth.features(3000, 8, seed=1)
    for kw in (dict(n_clusters=25), dict(dist_cutoff=0.7), dict(n_clusters=5, dist_cutoff=0.2)):
        a = kc.kcenters(X, "euclidean", **kw)
        b = oc.kcenters(X, od.euclidean, **kw)
        assert [int(i) for i in a.center_indices] == [int(i) for i in b.center_indices]
        assert_array_equal(a.assignments, b.assignments)
        assert_array_equal(a.distances, b.distances)
    T = od.Trajectory(frame0_xyz)
    a = kc.kcenters(T, "rmsd", dist_cutoff=0.1)
    b = oc.kcenters(T, od.rmsd, dist_cutoff=0.1)
    assert [int(i) for i in a.center_indices] == [int(i) for i in b.center_indices]
    assert_array_equal(a.distances, b.distances)


def test_pam_and_hybrid_restatement(ref, frame0_xyz):
    kc, km, hy, ut = ref[0], ref[1], ref[2], ref[3]
    from enspara_b200 import synth
    X = synth.features(1500, 6, seed=2)
    a = hy.hybrid(X, "euclidean", n_clusters=9, n_iters=3, random_state=4)
    b = oc.hybrid(X, od.euclidean, n_clusters=9, n_iters=3, random_state=4)
    assert [int(i) for i in a.center_indices] == [int(i) for i in b.center_indices]
    assert_array_equal(a.assignments, b.assignments)
    assert_array_equal(a.distances, b.distances)
    T = od.Trajectory(frame0_xyz)
    r = kc.kcenters(T, "rmsd", n_clusters=4)
    ia, da, aa, _ = km._kmedoids_pam_update(T, od.rmsd, list(r.center_indices),
                                            r.assignments.copy(), r.distances.copy(),
                                            random_state=7)
    ib, db, ab, _ = oc.pam_update(T, od.rmsd, list(r.center_indices), r.assignments.copy(),
                                  r.distances.copy(), random_state=7)
    assert [int(i) for i in ia] == [int(i) for i in ib]
    assert_array_equal(aa, ab)
    assert_array_equal(da, db)
    xa, xd = ut.assign_to_nearest_center(T, [T[i] for i in (3, 100, 250)], od.rmsd)
    ya, yd = oc.assign_to_nearest_center(T, [T[i] for i in (3, 100, 250)], od.rmsd)
    assert_array_equal(xa, ya)
    assert_array_equal(xd, yd)


def test_goldens_are_current(ref, golden, frame0_xyz):
    """tests/golden/reference_runs.npz is what the reference produces today."""
    kc = ref[0]
    from enspara_b200 import synth
    X = synth.features(5000, 16, seed=7)
    r = kc.kcenters(X, "euclidean", n_clusters=40)
    assert_array_equal(np.array(r.center_indices), golden["feat_k40_centers"])
    assert_array_equal(r.distances, golden["feat_k40_dist"])
    T = od.Trajectory(frame0_xyz)
    r = kc.kcenters(T, "rmsd", n_clusters=3)
    assert_array_equal(np.array(r.center_indices), golden["frame0_k3_centers"])
