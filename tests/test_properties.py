"""CPU property tests (hypothesis) for host-side logic: vectorised helpers against the
reference's loop semantics, and the HDF5 round trip for arbitrary shapes / dtypes."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from numpy.testing import assert_array_equal

from enspara_b200 import ra
from enspara_b200.util import h5min

_settings = settings(max_examples=60, deadline=None,
                     suppress_health_check=[HealthCheck.function_scoped_fixture])


@_settings
@given(st.integers(1, 60), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_find_cluster_centers_equals_the_reference_loop(n, k, seed):
    """cluster/util.py:208-242: for every label in np.unique order, the FIRST frame of minimum
    distance (np.where(...)[0][argmin])."""
    from enspara_b200.cluster import util
    rs = np.random.RandomState(seed)
    assignments = rs.randint(0, k, n)
    distances = rs.randint(0, 4, n).astype(float) / 2.0       # many exact ties
    got = util.find_cluster_centers(assignments, distances)
    want = []
    for label in np.unique(assignments):
        inds = np.where(assignments == label)[0]
        want.append(inds[np.argmin(distances[inds])])
    assert_array_equal(got, want)


@_settings
@given(st.lists(st.integers(1, 9), min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1))
def test_partition_indices_and_list_invert_concatenation(lengths, seed):
    """ra.py:223-242, 361-376: concatenated index <-> (trajectory, frame)."""
    rs = np.random.RandomState(seed)
    total = int(np.sum(lengths))
    idx = rs.randint(0, total, 7)
    pairs = ra.partition_indices(idx, lengths)
    starts = np.concatenate([[0], np.cumsum(lengths)])
    assert [int(starts[t] + f) for t, f in pairs] == [int(i) for i in idx]
    flat = np.arange(total)
    parts = ra.partition_list(flat, lengths)
    assert [len(p) for p in parts] == list(lengths)
    assert_array_equal(np.concatenate(parts), flat)


@_settings
@given(lengths=st.lists(st.integers(1, 7), min_size=2, max_size=6), size=st.integers(2, 3),
       seed=st.integers(0, 2 ** 31 - 1))
def test_ctr_ids_mpi_inverts_convert_local_indices(monkeypatch, lengths, size, seed):
    """kmedoids.py:365-408 and mpi/ops.py:14-39 are inverse maps when trajectory i lives on
    rank i % size (checked for an emulated world size, no process group needed)."""
    from enspara_b200 import mpi
    from enspara_b200.cluster import kmedoids
    monkeypatch.setattr(mpi, "size", lambda: size)
    rs = np.random.RandomState(seed)
    total = int(np.sum(lengths))
    glob = sorted(set(int(i) for i in rs.randint(0, total, 5)))
    pairs = kmedoids.ctr_ids_mpi(glob, lengths)
    back = mpi.ops.convert_local_indices(pairs, lengths)
    assert [int(b) for b in back] == glob


_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.float32, np.float64]


@_settings
@given(specs=st.lists(st.tuples(st.sampled_from(range(len(_DTYPES))),
                                st.lists(st.integers(0, 5), min_size=1, max_size=3)),
                      min_size=1, max_size=5), seed=st.integers(0, 2 ** 31 - 1))
def test_h5min_round_trip_any_shape_and_dtype(tmp_path, specs, seed):
    rs = np.random.RandomState(seed)
    arrays = {}
    for i, (di, shape) in enumerate(specs):
        arrays["node_%d" % i] = (rs.rand(*shape) * 100).astype(_DTYPES[di])
    p = os.path.join(str(tmp_path), "p.h5")
    h5min.write(p, arrays)
    f = h5min.File(p)
    assert sorted(f.keys()) == sorted(arrays)
    for k, v in arrays.items():
        got = h5min.read(p, k)
        assert got.shape == v.shape and got.dtype == v.dtype
        assert_array_equal(got, v)
