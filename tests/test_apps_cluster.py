"""`cluster` app: flag handling (CPU) and end-to-end feature clustering through the CLI (GPU),
after enspara/test/test_apps_cluster.py (feature inputs; trajectory inputs need mdtraj)."""
import os

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from enspara_b200.apps import cluster as app
from enspara_b200.exception import ImproperlyConfigured


def _argv(tmp, extra, feats=("a.npy", "b.npy")):
    return ["cluster", "--features"] + [os.path.join(tmp, f) for f in feats] + [
        "--distances", os.path.join(tmp, "dist.npy"),
        "--assignments", os.path.join(tmp, "assig.npy"),
        "--center-features", os.path.join(tmp, "ctrs.npy"),
        "--center-indices", os.path.join(tmp, "inds.npy")] + extra


def test_flag_validation(tmp_path):
    t = str(tmp_path)
    ok = ["--algorithm", "kcenters", "--cluster-distance", "euclidean", "--cluster-radius", "3"]
    args = app.process_command_line(_argv(t, ok))
    assert args.Clusterer.__name__ == "KCenters" and args.cluster_radius == 3.0
    with pytest.raises(ImproperlyConfigured):   # no radius / number
        app.process_command_line(_argv(t, ["--algorithm", "kcenters", "--cluster-distance",
                                           "euclidean"]))
    with pytest.raises(ImproperlyConfigured):   # rmsd on features
        app.process_command_line(_argv(t, ["--algorithm", "kcenters", "--cluster-distance",
                                           "rmsd", "--cluster-number", "3"]))
    with pytest.raises(ImproperlyConfigured):   # iterations with kcenters
        app.process_command_line(_argv(t, ok + ["--cluster-iterations", "2"]))
    with pytest.raises(ImproperlyConfigured):   # radius with kmedoids
        app.process_command_line(_argv(t, ["--algorithm", "kmedoids", "--cluster-distance",
                                           "euclidean", "--cluster-radius", "3"]))
    with pytest.raises(ImproperlyConfigured):   # restart flags only for kmedoids
        app.process_command_line(_argv(t, ok + ["--init-center-inds", "x.npy"]))
    with pytest.raises(ImproperlyConfigured):   # --atoms with features
        app.process_command_line(_argv(t, ok + ["--atoms", "name CA"]))
    with pytest.raises(SystemExit):             # unknown algorithm
        app.process_command_line(_argv(t, ["--algorithm", "dbscan", "--cluster-number", "3"]))


@pytest.mark.gpu
def test_feature_clustering_end_to_end(tmp_path):
    """test_apps_cluster.py:366-387 (blobs seed 3, euclidean, radius 3 -> 11 clusters) and
    :509-548 (k-hybrid with 0 sweeps == KCenters)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sklearn.datasets import make_blobs
    from enspara_b200 import ra
    from enspara_b200.cluster import KCenters
    t = str(tmp_path)
    X, _ = make_blobs(n_samples=100, n_features=3, random_state=3)
    a, b = X[:50], X[50:]
    np.save(os.path.join(t, "a.npy"), a)
    np.save(os.path.join(t, "b.npy"), b)
    rc = app.main(_argv(t, ["--algorithm", "kcenters", "--cluster-distance", "euclidean",
                            "--cluster-radius", "3"]))
    assert rc == 0
    assig = ra.load(os.path.join(t, "assig.npy"))
    dist = ra.load(os.path.join(t, "dist.npy"))
    inds = np.load(os.path.join(t, "inds.npy"))
    ctrs = np.load(os.path.join(t, "ctrs.npy"))
    assert assig.shape == (2, 50) and dist.shape == (2, 50)
    assert assig.dtype == np.int64 and dist.dtype == np.float64
    assert len(np.unique(assig)) == 11 and len(inds) == 11 and ctrs.shape == (11, 3)
    assert dist.max() < 3
    direct = KCenters("euclidean", cluster_radius=3).fit(X)
    assert_array_equal(assig.reshape(-1), direct.labels_)
    assert_array_equal(dist.reshape(-1), direct.distances_)
    assert_array_equal(ctrs, X[[int(i) for i in direct.center_indices_]])
    assert_array_equal(inds, [(int(i) // 50, int(i) % 50) for i in direct.center_indices_])

    rc = app.main(_argv(t, ["--algorithm", "khybrid", "--cluster-distance", "euclidean",
                            "--cluster-number", "3", "--cluster-iterations", "0"]))
    assert rc == 0
    d0 = ra.load(os.path.join(t, "dist.npy")).reshape(-1)
    assert_array_equal(d0, KCenters("euclidean", n_clusters=3).fit(X).distances_)

    # ragged inputs -> RaggedArray outputs (.npz next to the requested name)
    np.save(os.path.join(t, "c.npy"), X[:30])
    np.save(os.path.join(t, "d.npy"), X[30:])
    rc = app.main(_argv(t, ["--algorithm", "khybrid", "--cluster-distance", "manhattan",
                            "--cluster-number", "4", "--cluster-iterations", "1"],
                        feats=("c.npy", "d.npy")))
    assert rc == 0
    r = ra.load(os.path.join(t, "assig.npy.npz"))
    assert [len(x) for x in r] == [30, 70] and len(np.unique(r.flatten())) == 4


@pytest.mark.gpu
def test_h5_io_intermediates_and_kmedoids_restart(tmp_path):
    """.h5 outputs / inputs without PyTables (test_apps_cluster.py:26-92 checks these files with
    ra.load), --save_intermediates (hybrid.py:129-151, kmedoids.py:459-473) and the k-medoids
    restart flags (apps/cluster.py:136-147, 317-329)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sklearn.datasets import make_blobs
    from enspara_b200 import ra
    from enspara_b200.cluster import KHybrid, KMedoids
    t = str(tmp_path)
    X, _ = make_blobs(n_samples=120, n_features=4, random_state=5)
    rows = ra.RaggedArray([X[:50], X[50:]])
    ra.save(os.path.join(t, "feat.h5"), rows)                 # ragged features in one .h5
    out = ["--distances", os.path.join(t, "dist.h5"), "--assignments", os.path.join(t, "assig.h5"),
           "--center-features", os.path.join(t, "ctrs.npy"),
           "--center-indices", os.path.join(t, "inds.npy")]
    rc = app.main(["cluster", "--features", os.path.join(t, "feat.h5")] + out + [
        "--algorithm", "khybrid", "--cluster-distance", "euclidean", "--cluster-number", "5",
        "--cluster-iterations", "2", "--save_intermediates", "True"])
    assert rc == 0
    assig = ra.load(os.path.join(t, "assig.h5"))
    dist = ra.load(os.path.join(t, "dist.h5"))
    assert isinstance(assig, ra.RaggedArray) and list(assig.lengths) == [50, 70]
    assert assig.dtype == np.int64 and dist.dtype == np.float64
    direct = KHybrid("euclidean", n_clusters=5, kmedoids_updates=2).fit(X)
    assert len(np.unique(assig.flatten())) == 5
    assert dist.flatten().shape == direct.distances_.shape
    # intermediates: after k-centers and after the first of the two sweeps
    for tag in ("kcenters", "kmedoids-0"):
        d = os.path.join(t, "intermediate-%s" % tag)
        assert sorted(os.listdir(d)) == ["assig.h5", "ctrs.npy", "dist.h5", "inds.npy"], tag
        assert list(ra.load(os.path.join(d, "assig.h5")).lengths) == [50, 70]
    # the k-centers intermediate equals a plain k-centers run
    from enspara_b200.cluster import KCenters
    kc = KCenters("euclidean", n_clusters=5).fit(X)
    assert_array_equal(ra.load(os.path.join(t, "intermediate-kcenters", "dist.h5")).flatten(),
                       kc.distances_)

    # restart k-medoids from these results (files written by the run above)
    np.save(os.path.join(t, "a.npy"), X[:50])
    np.save(os.path.join(t, "b.npy"), X[50:])
    # ... exactly as a user would: the previous run's .h5 / .npy outputs are the init files
    os.replace(os.path.join(t, "assig.h5"), os.path.join(t, "init_assig.h5"))
    os.replace(os.path.join(t, "dist.h5"), os.path.join(t, "init_dist.h5"))
    os.replace(os.path.join(t, "inds.npy"), os.path.join(t, "init_inds.npy"))
    rc = app.main(_argv(t, ["--algorithm", "kmedoids", "--cluster-distance", "euclidean",
                            "--cluster-number", "5", "--cluster-iterations", "1",
                            "--init-assignments", os.path.join(t, "init_assig.h5"),
                            "--init-distances", os.path.join(t, "init_dist.h5"),
                            "--init-center-inds", os.path.join(t, "init_inds.npy")]))
    assert rc == 0
    d2 = ra.load(os.path.join(t, "dist.npy.npz")).flatten()
    # a PAM sweep never increases the mean-square cost
    assert np.mean(d2 ** 2) <= np.mean(dist.flatten() ** 2) + 1e-12


def test_main_dispatcher_routes_to_the_apps(tmp_path, monkeypatch):
    """apps/main.py:6-62: `enspara cluster ...` / `enspara reassign ...`."""
    from enspara_b200.apps import main as entry
    from enspara_b200.apps import reassign as rapp
    seen = {}
    monkeypatch.setattr(app, "main", lambda argv: seen.setdefault("cluster", argv) and 0)
    monkeypatch.setattr(rapp, "main", lambda argv: seen.setdefault("reassign", argv) and 0)
    assert entry.main(["enspara", "cluster", "--algorithm", "kcenters"]) == 0
    assert seen["cluster"] == ["cluster", "--algorithm", "kcenters"]
    assert entry.main(["enspara", "reassign", "--centers", "c.pkl"]) == 0
    assert seen["reassign"] == ["reassign", "--centers", "c.pkl"]
    with pytest.raises(SystemExit):
        entry.main(["enspara", "implied"])
