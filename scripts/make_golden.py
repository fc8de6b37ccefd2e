"""Generate tests/golden/* from the reference (run in the build container only).

  frame0_xyz.npy        the reference's RMSD test fixture enspara/test/data/frame0.xtc decoded
                        with oracle/xtc.py (501 frames x 22 atoms, float32 nm)
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

    kc, km, hy, ut, libdist, mpi = rh.modules()
    out = {}

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
