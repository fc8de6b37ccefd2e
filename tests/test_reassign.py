"""`reassign` / `batch_reassign` and the reassign app, after
enspara/test/test_apps_reassign.py.  CPU: batching rule, flag validation, file access.
GPU: streamed re-assignment over files == assign_to_nearest_center on the concatenation ==
the oracle."""
import os
import pickle

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from enspara_b200 import ra
from enspara_b200.apps import reassign as app
from enspara_b200.cluster import reassign as rz
from enspara_b200.exception import ImproperlyConfigured


def test_compute_batches_follows_the_reference_rule():
    # util.py:551-567: greedy, strict '<' against batch_size
    assert rz.compute_batches([5, 5, 5], 11) == [[0, 1], [2]]
    assert rz.compute_batches([5, 5, 5], 10) == [[0], [1], [2]]
    assert rz.compute_batches([3, 9, 1, 1], 10) == [[0], [1], [2, 3]]
    assert rz.compute_batches([], 10) == [[]]
    assert rz.compute_batches([12], 10) == [[], [0]]


def test_first_file_filling_a_batch_leaves_no_empty_batch():
    # lengths[0] >= batch_size: compute_batches opens with [] (reference rule); the streaming
    # pass must not see it (it chains a prefetch from every batch to the next)
    assert rz.compute_batches([100, 10, 20], 100) == [[], [0], [1, 2]]
    assert rz.nonempty_batches([100, 10, 20], 100) == [[0], [1, 2]]
    assert rz.nonempty_batches([], 10) == []
    assert rz.nonempty_batches([5, 5, 5], 11) == [[0, 1], [2]]


def test_sound_and_load(tmp_path, frame0_h5_xyz):
    from enspara_b200.util import h5min
    p_npy = str(tmp_path / "a.npy")
    p_h5 = str(tmp_path / "a.h5")
    np.save(p_npy, frame0_h5_xyz[:40])
    h5min.write(p_h5, {"coordinates": frame0_h5_xyz[:30]})
    assert rz.sound_trajectory(p_npy) == 40
    assert rz.sound_trajectory(p_h5) == 30
    assert_array_equal(rz.load_frames(p_h5), frame0_h5_xyz[:30])
    sub = rz.load_frames(p_npy, atom_indices=[0, 3, 5])
    assert_array_equal(sub, frame0_h5_xyz[:40][:, [0, 3, 5]])
    with pytest.raises(ImproperlyConfigured):      # a format only mdtraj reads
        rz.sound_trajectory(str(tmp_path / "a.dcd"))
    from enspara_b200.exception import DataInvalid
    with pytest.raises(DataInvalid):                # .xtc is read natively: a missing file
        rz.sound_trajectory(str(tmp_path / "a.xtc"))


def test_app_flag_validation(tmp_path):
    t = str(tmp_path)
    trj = os.path.join(t, "a.npy")
    np.save(trj, np.zeros((3, 4, 3), np.float32))
    base = ["reassign", "--centers", os.path.join(t, "c.npy"), "--distances",
            os.path.join(t, "d.h5"), "--assignments", os.path.join(t, "a.h5")]
    args = app.process_command_line(base + ["--trajectories", trj, "--topology", "x.pdb"])
    assert args.output_path == t and args.mem_fraction == 0.5
    with pytest.raises(ImproperlyConfigured):
        app.process_command_line(base + ["--trajectories", trj, "--topology", "x.pdb",
                                         "-m", "1.5"])
    with pytest.raises(ImproperlyConfigured):
        app.process_command_line(base + ["--trajectories", trj, "--topology", "x.pdb",
                                         "--topology", "y.pdb"])
    with pytest.raises(FileNotFoundError):
        app.process_command_line(base + ["--trajectories", os.path.join(t, "nope.npy"),
                                         "--topology", "x.pdb"])


@pytest.mark.gpu
def test_reassign_streams_files_and_matches_oracle(tmp_path, frame0_xyz, monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from enspara_b200.cluster import util
    from oracle import cluster as oc
    from oracle import distances as od
    t = str(tmp_path)
    cuts = [0, 120, 121, 300, 501]                     # ragged lengths incl. a 1-frame file
    files = []
    for i in range(4):
        p = os.path.join(t, "trj%d.npy" % i)
        np.save(p, frame0_xyz[cuts[i]:cuts[i + 1]])
        files.append(p)
    centers = frame0_xyz[::50]
    # force several batches: at most 201 frames per batch -> [[0, 1], [2], [3]]
    monkeypatch.setattr(rz, "determine_batch_size", lambda *a, **k: (202, 0.0))
    stats = {}
    lengths = [rz.sound_trajectory(f) for f in files]
    a_list, d_list = rz.batch_reassign([(f, None, None) for f in files], centers, lengths,
                                       0.5, stats=stats)
    assert stats["batches"] == 3 and [len(a) for a in a_list] == lengths
    ea, ed = util.assign_to_nearest_center(od.Trajectory(frame0_xyz), centers, "rmsd")
    assert_array_equal(np.concatenate(a_list), ea)
    assert_array_equal(np.concatenate(d_list), ed)
    oa, odist = oc.assign_to_nearest_center(
        od.Trajectory(frame0_xyz), [od.Trajectory(c[None]) for c in centers], od.rmsd)
    assert_array_equal(np.concatenate(a_list), oa)
    assert_allclose(np.concatenate(d_list), odist, rtol=1e-5, atol=1e-6)

    # the function and the app: ragged files -> RaggedArray outputs in the reference's .h5 layout
    assig, dist = rz.reassign(["top.pdb"], [files], ["all"], centers)
    assert isinstance(assig, ra.RaggedArray) and list(assig.lengths) == lengths
    with open(os.path.join(t, "ctrs.pkl"), "wb") as f:
        pickle.dump([od.Trajectory(c[None]) for c in centers], f)
    rc = app.main(["reassign", "--centers", os.path.join(t, "ctrs.pkl"), "--trajectories",
                   files[0], files[1], "--trajectories", files[2], files[3],
                   "--topology", "a.pdb", "--topology", "b.pdb", "--atoms", "all",
                   "--distances", os.path.join(t, "dist.h5"),
                   "--assignments", os.path.join(t, "assig.h5")])
    assert rc == 0
    back = ra.load(os.path.join(t, "assig.h5"))
    assert list(back.lengths) == lengths and back.dtype == np.int64
    assert_array_equal(back.flatten(), ea)
    assert_array_equal(ra.load(os.path.join(t, "dist.h5")).flatten(), ed)

    # equal lengths -> plain 2-D arrays (util.py:724-729)
    np.save(os.path.join(t, "e0.npy"), frame0_xyz[:100])
    np.save(os.path.join(t, "e1.npy"), frame0_xyz[100:200])
    assig, dist = rz.reassign(["t"], [[os.path.join(t, "e0.npy"), os.path.join(t, "e1.npy")]],
                              ["all"], centers)
    assert isinstance(assig, np.ndarray) and assig.shape == (2, 100)
    assert_array_equal(assig.reshape(-1), ea[:200])

    # the first file exactly fills a batch (lengths[0] == batch_size; the case that used to
    # leave the prefetch chain without a head)
    first_big = [files[2], files[0], files[1]]          # 179, 120, 1 frames
    monkeypatch.setattr(rz, "determine_batch_size", lambda *a, **k: (179, 0.0))
    l2 = [rz.sound_trajectory(f) for f in first_big]
    a2, d2 = rz.batch_reassign([(f, None, None) for f in first_big], centers, l2, 0.5)
    assert [len(a) for a in a2] == l2
    assert_array_equal(np.concatenate(a2), np.concatenate([ea[121:300], ea[:120], ea[120:121]]))

    with pytest.raises(ImproperlyConfigured):          # batch smaller than the largest file
        monkeypatch.setattr(rz, "determine_batch_size", lambda *a, **k: (100, 0.0))
        rz.batch_reassign([(f, None, None) for f in files], centers, lengths, 0.5)


def test_determine_batch_size_bounds(monkeypatch):
    """Two pinned staging buffers + the device copy bound a batch; the cap never cuts below the
    largest file while that still fits the hard (RAM / HBM) bound."""
    import types
    import psutil
    import torch
    monkeypatch.setattr(psutil, "virtual_memory",
                        lambda: types.SimpleNamespace(total=64 << 30))
    # the RAM bound alone, whether or not the box running this has a GPU
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    bpf = 500 * 3 * 4
    hard = int((64 << 30) * 0.5 / 2 / bpf)
    n, gb = rz.determine_batch_size(500, 4, 0.5)
    assert n == min(hard, rz.BATCH_BYTES_CAP // bpf) and abs(gb - n * bpf / 2 ** 30) < 1e-9
    big = rz.BATCH_BYTES_CAP // bpf + 1000              # a file larger than the soft cap
    n2, _ = rz.determine_batch_size(500, 4, 0.5, largest_file=big)
    assert n2 == big
    n3, _ = rz.determine_batch_size(500, 4, 0.5, largest_file=10 * hard)
    assert n3 == hard                                   # ... but never beyond the hard bound
    # with a device: a quarter of the free HBM over 3x the host bytes per frame bounds it too
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda: (24 << 30, 180 << 30))
    n4, _ = rz.determine_batch_size(500, 4, 0.5, largest_file=10 * hard)
    assert n4 == int((24 << 30) * 0.25 / (3 * bpf)) < hard
