"""Trajectory input without mdtraj (the loaders either side of the path):
the native .xtc decoder of libenspara_b200.so (csrc/eb_xtc.cu), the .pdb topology with the
atom-selection mini-language (enspara_b200/util/traj.py), and the `cluster` / `reassign` apps
running on the reference's own trajectory fixture -- tests/golden/frame0.xtc + native.pdb are
copies of enspara/test/data/frame0.xtc / native.pdb (test DATA of the reference; its CLI tests
enspara/test/test_apps_cluster.py:97-210 run on exactly these files).

CPU: decoder vs the golden coordinates (decoded independently by oracle/xtc.py, a pure-Python
reader written from the published format), selections, loaders, error behaviour.
GPU: the CLI end to end, like test_apps_cluster.py::test_rmsd_cluster_basic* and
::test_rmsd_cluster_selection / ::test_rmsd_cluster_broken_atoms / subsample + reassign.
"""
import os
import pickle

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from enspara_b200.exception import ImproperlyConfigured
from enspara_b200.util import traj

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
XTC = os.path.join(GOLD, "frame0.xtc")
PDB = os.path.join(GOLD, "native.pdb")
SEL = "(name N or name C or name CA or name H or name O)"


def test_native_xtc_decoder_matches_golden_coordinates(frame0_xyz):
    assert traj.xtc_shape(XTC) == (501, 22)
    assert_array_equal(traj.read_xtc(XTC), frame0_xyz)                 # bit for bit
    assert_array_equal(traj.read_xtc(XTC, stride=7), frame0_xyz[::7])
    assert_array_equal(traj.read_xtc(XTC, stride=3, first=5), frame0_xyz[5::3])
    idx = [8, 4, 21, 0]                                                # order is kept
    assert_array_equal(traj.read_xtc(XTC, stride=2, atom_indices=idx), frame0_xyz[::2][:, idx])
    assert_array_equal(traj.read_xtc(XTC, first=500), frame0_xyz[500:])
    assert traj.read_xtc(XTC, first=501).shape == (0, 22, 3)
    assert_array_equal(traj.read_xtc(XTC, first=17, max_frames=1), frame0_xyz[17:18])
    out = np.empty((167, 22, 3), np.float32)
    assert traj.read_xtc(XTC, stride=3, out=out) is out
    assert_array_equal(out, frame0_xyz[::3])


def test_xtc_errors(tmp_path):
    from enspara_b200.exception import DataInvalid
    bad = tmp_path / "bad.xtc"
    bad.write_bytes(b"\x00" * 100)
    with pytest.raises(DataInvalid):
        traj.xtc_shape(str(bad))
    with pytest.raises(DataInvalid):
        traj.xtc_shape(str(tmp_path / "missing.xtc"))
    trunc = tmp_path / "trunc.xtc"
    trunc.write_bytes(open(XTC, "rb").read()[:5000])
    with pytest.raises(DataInvalid):
        traj.read_xtc(str(trunc))
    with pytest.raises(DataInvalid):
        traj.read_xtc(XTC, atom_indices=[0, 22])                       # out of range


def test_pdb_topology_and_selections():
    top = traj.load_pdb_topology(PDB)
    assert top.n_atoms == 22 and list(top.resnames[[0, 8, 21]]) == ["ACE", "ALA", "NME"]
    assert_array_equal(top.select("all"), np.arange(22))
    # the selections of the reference's CLI tests (test_apps_cluster.py:103, 160, 238-241)
    assert_array_equal(top.select(SEL), [4, 5, 6, 7, 8, 14, 15, 16, 17])
    assert_array_equal(top.select("(name N or name C or name CA)"), [4, 6, 8, 14, 16])
    assert_array_equal(top.select("(name N or name O) and (residue 2)"), [6, 15])
    assert_array_equal(top.select("(name CA) and (residue 2 or residue 3)"), [8])
    assert_array_equal(top.select("backbone"), [4, 5, 6, 8, 14, 15, 16])
    assert_array_equal(top.select("not element H"), [1, 4, 5, 6, 8, 10, 14, 15, 16, 18])
    assert_array_equal(top.select("index 0 to 4"), np.arange(5))
    assert_array_equal(top.select("resid 1 and name CB"), [10])
    assert_array_equal(top.select("residue >= 3"), np.arange(16, 22))
    assert_array_equal(top.select("resname ALA and not (name N or name CA)"),
                       [7, 9, 10, 11, 12, 13, 14, 15])
    for bad in ("residue -1", "foo 3", "name", "(name N", "name N)", ""):
        with pytest.raises(ValueError):
            top.select(bad)
    sub = top.subset(top.select(SEL))
    assert sub.n_atoms == 9 and list(sub.names) == ["C", "O", "N", "H", "CA", "C", "O", "N", "H"]


def test_load_like_mdtraj(frame0_xyz):
    top = traj.load_pdb_topology(PDB)
    idx = top.select(SEL)
    t = traj.load(XTC, top=PDB, stride=4, atom_indices=idx)
    assert len(t) == 126 and t.n_atoms == 9 and t.top.n_atoms == 9
    assert_array_equal(t.xyz, frame0_xyz[::4][:, idx])
    assert_array_equal(t[3].xyz, frame0_xyz[12:13][:, idx])            # int -> 1-frame object
    assert_array_equal(t[[1, 5]].xyz, frame0_xyz[[4, 20]][:, idx])
    mask = np.zeros(126, bool)
    mask[[2, 7]] = True
    assert len(t[mask]) == 2
    again = type(t)(xyz=t.xyz[:2], topology=t.top)                     # mpi/ops.py:208-210
    assert len(again) == 2
    f = traj.load_frame(XTC, 40, top=PDB)
    assert f.n_atoms == 22
    assert_array_equal(f.xyz[0], frame0_xyz[40])
    p = traj.load(PDB)
    assert p.xyz.shape == (1, 22, 3) and abs(float(p.xyz[0, 0, 0]) - 0.43) < 1e-6   # A -> nm


def test_cluster_io_loads_trajectories_without_mdtraj(frame0_xyz):
    from enspara_b200.cluster import io as cio
    lengths, data = cio.load_trajectories([PDB], [[XTC, XTC]], [SEL], 2)
    idx = traj.load_pdb_topology(PDB).select(SEL)
    assert list(lengths) == [251, 251] and data.xyz.shape == (502, 9, 3)
    assert_array_equal(data.xyz[:251], frame0_xyz[::2][:, idx])
    assert data.top.n_atoms == 9
    with pytest.raises(ImproperlyConfigured):
        cio.load_trajectories([PDB], [[XTC]], ["residue -1"], 1)
    with pytest.raises(ImproperlyConfigured):
        cio.load_trajectories([PDB], [[XTC]], ["name ZZ"], 1)          # selects nothing
    from enspara_b200.cluster import reassign as rz
    assert rz.sound_trajectory(XTC) == 501 and rz.sound_trajectory(XTC, stride=4) == 126
    top, ids = rz._select_atoms(PDB, SEL)
    assert_array_equal(ids, idx)
    out = np.zeros((501, 9, 3), np.float32)
    rz.load_frames(XTC, top, ids, out=out)
    assert_array_equal(out, frame0_xyz[:, idx])


def _run_cli(tmp, extra, trajs=(XTC, XTC)):
    from enspara_b200.apps import cluster as app
    out = {k: os.path.join(tmp, v) for k, v in (("d", "dist.h5"), ("a", "assig.h5"),
                                                ("c", "ctrs.pkl"), ("i", "inds.npy"))}
    argv = ["cluster", "--trajectories"] + list(trajs) + ["--topology", PDB] + extra + [
        "--distances", out["d"], "--assignments", out["a"], "--center-features", out["c"],
        "--center-indices", out["i"]]
    assert app.main(argv) == 0
    return out


@pytest.mark.gpu
def test_cli_rmsd_clustering_on_the_reference_fixture(tmp_path, frame0_xyz):
    """test_apps_cluster.py:97-135: two copies of frame0.xtc, the five backbone-ish names,
    k-hybrid / k-centers with a radius or a fixed k; outputs of shape (2, 501), int / float."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import ra
    from enspara_b200.cluster import KCenters
    idx = traj.load_pdb_topology(PDB).select(SEL)
    X = np.concatenate([frame0_xyz[:, idx], frame0_xyz[:, idx]])
    t = str(tmp_path)
    out = _run_cli(t, ["--cluster-radius", "0.1", "--atoms", SEL, "--algorithm", "kcenters"])
    a, d = ra.load(out["a"]), ra.load(out["d"])
    assert a.shape == (2, 501) and a.dtype == np.int64 and d.shape == (2, 501)
    want = KCenters("rmsd", cluster_radius=0.1).fit(X).result_
    assert_array_equal(np.asarray(a).reshape(-1), want.assignments)
    assert_array_equal(np.asarray(d).reshape(-1), want.distances)
    assert float(np.asarray(d).max()) < 0.1
    inds = np.load(out["i"])
    assert [tuple(r) for r in inds] == [(int(c) // 501, int(c) % 501)
                                        for c in want.center_indices]
    # centres: FULL-topology structures re-loaded from the files (util.py:407-431, 505-508)
    with open(out["c"], "rb") as fh:
        ctrs = pickle.load(fh)
    assert len(ctrs) == len(inds) and all(c.n_atoms == 22 and len(c) == 1 for c in ctrs)
    for c, (ti, fi) in zip(ctrs, inds):
        assert_array_equal(c.xyz[0], frame0_xyz[fi])
    # fixed k (test_apps_cluster.py:124-135) and k-hybrid (:97-108)
    out = _run_cli(t, ["--cluster-number", "10", "--atoms", SEL, "--algorithm", "kcenters"])
    assert_array_equal(np.unique(np.asarray(ra.load(out["a"]))), np.arange(10))
    out = _run_cli(t, ["--cluster-radius", "0.1", "--atoms", SEL, "--algorithm", "khybrid"])
    assert ra.load(out["a"]).shape == (2, 501)
    # another selection (test_apps_cluster.py:154-166) and the broken one (:138-151)
    out = _run_cli(t, ["--cluster-radius", "0.1", "--atoms", "(name N or name C or name CA)",
                       "--algorithm", "khybrid"])
    assert ra.load(out["d"]).shape == (2, 501)
    with pytest.raises(ImproperlyConfigured):
        _run_cli(t, ["--cluster-radius", "0.1", "--atoms", "residue -1", "--algorithm",
                     "khybrid"])


@pytest.mark.gpu
def test_cli_subsample_and_reassign_on_xtc(tmp_path, frame0_xyz):
    """test_apps_cluster.py:169-210: --subsample clusters every 4th frame, then every frame of
    the input files is re-assigned (streamed through cluster/reassign.py, native .xtc loader);
    --no-reassign writes no assignments."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200 import ra
    from enspara_b200.cluster import util
    idx = traj.load_pdb_topology(PDB).select(SEL)
    t = str(tmp_path)
    out = _run_cli(t, ["--cluster-radius", "0.1", "--atoms", SEL, "--algorithm", "kcenters",
                       "--subsample", "4"])
    a, d = ra.load(out["a"]), ra.load(out["d"])
    assert np.asarray(a).shape == (2, 501)
    inds = np.load(out["i"])
    assert all(int(f) % 4 == 0 for _, f in inds)                   # frame * subsample
    with open(out["c"], "rb") as fh:
        ctrs = pickle.load(fh)
    for c, (ti, fi) in zip(ctrs, inds):
        assert_array_equal(c.xyz[0], frame0_xyz[fi])                   # full structures
    centers = [frame0_xyz[int(f)][idx] for _, f in inds]
    ea, ed = util.assign_to_nearest_center(frame0_xyz[:, idx], centers, "rmsd")
    assert_array_equal(np.asarray(a)[0], ea)
    assert_array_equal(np.asarray(d)[1], ed)
    for f in (out["a"], out["d"]):
        os.remove(f)
    out = _run_cli(t, ["--cluster-radius", "0.1", "--atoms", SEL, "--algorithm", "kcenters",
                       "--subsample", "4", "--no-reassign"])
    assert not os.path.exists(out["a"]) and not os.path.exists(out["d"])
