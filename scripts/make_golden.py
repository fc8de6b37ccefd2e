"""Generate tests/golden/* from the reference (run in the build container only).

  frame0_xyz.npy        the reference's RMSD test fixture enspara/test/data/frame0.xtc decoded
                        with oracle/xtc.py (501 frames x 22 atoms, float32 nm)
  frame0_h5_xyz.npy     enspara/test/data/frame0.h5 ('/coordinates', 501 x 22 x 3 float32; the same
                        trajectory at 1e-3 nm precision) decoded with enspara_b200/util/h5min.py
                        -- the fixture of the reference's PAM goldens (test_cluster.py:533-554,
                        :378-419)
  reference_runs.npz    outputs of the REAL reference package (imported from /root/reference
                        through oracle/refharness.py: its own Python loops + its own Cython
                        libdist; mdtraj.rmsd replaced by the restated RMSD) on seeded inputs.

Usage: python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    import logging
    logging.disable(logging.CRITICAL)
    from sklearn.datasets import make_blobs

    from enspara_b200 import synth
    from oracle import distances as od
    from oracle import refharness as rh
    from oracle import xtc

    os.makedirs(GOLDEN, exist_ok=True)
    xyz, *_ = xtc.read_xtc(os.path.join(rh.REFERENCE, "enspara/test/data/frame0.xtc"))
    np.save(os.path.join(GOLDEN, "frame0_xyz.npy"), xyz)

    from enspara_b200.util import h5min
    xyz_h5 = h5min.read(os.path.join(rh.REFERENCE, "enspara/test/data/frame0.h5"),
                        "coordinates")
    assert xyz_h5.shape == (501, 22, 3) and xyz_h5.dtype == np.float32
    np.save(os.path.join(GOLDEN, "frame0_h5_xyz.npy"), xyz_h5)

    kc, km, hy, ut, libdist, mpi = rh.modules()
    out = {}

    # ---- the reference's PAM goldens on frame0.h5 (reference loops + restated md.rmsd) ----
    Th = od.Trajectory(xyz_h5)
    r = kc.kcenters(Th, od.rmsd, n_clusters=3)
    ind, d, a, _ = km._kmedoids_pam_update(Th, od.rmsd, list(r.center_indices),
                                           r.assignments.copy(), r.distances.copy(),
                                           random_state=0)
    assert [int(i) for i in ind] == [298, 44, 341], ind   # enspara/test/test_cluster.py:548
    out["h5_k3_pam_centers"] = np.array(ind)
    out["h5_k3_pam_assign"] = a
    out["h5_k3_pam_dist"] = d
    r = kc.kcenters(Th, od.rmsd, n_clusters=10)
    props = [int(np.where(r.assignments == cid)[0][0]) for cid in range(10)]
    ind, d, a, _ = km._kmedoids_pam_update(Th, od.rmsd, list(r.center_indices),
                                           r.assignments.copy(), r.distances.copy(),
                                           proposals=props, random_state=0)
    # enspara/test/test_cluster.py:410-411 (the MPI test; a single rank sees global indices)
    assert [int(i) for i in ind] == [0, 37, 400, 105, 12, 327, 242, 346, 42, 3], ind
    out["h5_k10_pam_proposals"] = np.array(props)
    out["h5_k10_pam_centers"] = np.array(ind)
    out["h5_k10_pam_assign"] = a
    out["h5_k10_pam_dist"] = d

    # ---- RMSD path on the reference's fixture (reference loops + restated md.rmsd) --------
    T = od.Trajectory(xyz)
    r = kc.kcenters(T, "rmsd", dist_cutoff=0.1)
    out["frame0_cut01_centers"] = np.array(r.center_indices)
    out["frame0_cut01_assign"] = r.assignments
    out["frame0_cut01_dist"] = r.distances
    r = kc.kcenters(T, "rmsd", n_clusters=3)
    out["frame0_k3_centers"] = np.array(r.center_indices)
    out["frame0_k3_assign"] = r.assignments
    out["frame0_k3_dist"] = r.distances
    ind, d, a, _ = km._kmedoids_pam_update(T, od.rmsd, list(r.center_indices),
                                           r.assignments.copy(), r.distances.copy(),
                                           random_state=0)
    out["frame0_k3_pam_centers"] = np.array(ind)
    out["frame0_k3_pam_assign"] = a
    out["frame0_k3_pam_dist"] = d
    r = hy.hybrid(T, "rmsd", n_clusters=5, n_iters=5, random_state=0)
    out["frame0_hybrid5_centers"] = np.array(r.center_indices)
    out["frame0_hybrid5_assign"] = r.assignments
    out["frame0_hybrid5_dist"] = r.distances

    # ---- euclidean path: the reference end to end (its own libdist) ------------------------
    X = synth.features(5000, 16, seed=7)
    r = kc.kcenters(X, "euclidean", n_clusters=40)
    out["feat_k40_centers"] = np.array(r.center_indices)
    out["feat_k40_assign"] = r.assignments
    out["feat_k40_dist"] = r.distances
    r = kc.kcenters(X, "euclidean", dist_cutoff=0.9)
    out["feat_cut09_centers"] = np.array(r.center_indices)
    out["feat_cut09_dist"] = r.distances
    r = kc.kcenters(X, "manhattan", n_clusters=25)
    out["feat_manh_k25_centers"] = np.array(r.center_indices)
    out["feat_manh_k25_assign"] = r.assignments
    out["feat_manh_k25_dist"] = r.distances
    r = hy.hybrid(X, "euclidean", n_clusters=12, n_iters=3, random_state=5)
    out["feat_hybrid12_centers"] = np.array(r.center_indices)
    out["feat_hybrid12_assign"] = r.assignments
    out["feat_hybrid12_dist"] = r.distances
    a, d = ut.assign_to_nearest_center(X, X[[5, 17, 99, 1234, 4000]], libdist.euclidean)
    out["feat_assign5_assign"] = a
    out["feat_assign5_dist"] = d
    r = km.kmedoids(X[:600], "euclidean", n_clusters=6, n_iters=4, random_state=3)
    out["feat_kmedoids6_centers"] = np.array(r.center_indices)
    out["feat_kmedoids6_assign"] = r.assignments
    out["feat_kmedoids6_dist"] = r.distances

    # ---- the reference's own blob goldens ---------------------------------------------------
    Xb, _ = make_blobs(centers=[(0, 0), (0, 10), (10, 0)], random_state=0)

    def sq(X, x):
        return np.square(X - x).sum(axis=1)
    r = kc.kcenters(Xb, sq, n_clusters=3)
    ind, d, a, _ = km._kmedoids_pam_update(Xb, sq, r.center_indices, r.assignments,
                                           r.distances, random_state=0)
    assert list(ind) == [0, 7, 17], ind      # enspara/test/test_cluster.py:507-530
    out["blobs_X"] = Xb
    out["blobs_pam_centers"] = np.array(ind)
    out["blobs_pam_assign"] = a
    out["blobs_pam_dist"] = d

    np.savez_compressed(os.path.join(GOLDEN, "reference_runs.npz"), **out)
    for k in sorted(out):
        print("%-28s %s %s" % (k, out[k].dtype, out[k].shape))


if __name__ == "__main__":
    main()
