"""CPU: the oracle pieces added for the CPU baselines -- they must not change what is
measured or compared:
  * the SSE restatement of mdtraj's float32 RMSD is BIT-IDENTICAL to the scalar lane emulation
    (so the timed baseline and the noise-quantification variant are the same arithmetic);
  * the C generator of the synthetic inputs is bit-identical to enspara_b200/synth.py (numpy),
    so the CPU arms need neither numpy's minutes nor the product's CUDA library;
  * the reference's own compiled libdist (oracle/_ref) loads without the reference's Python
    package and equals the C restatement (what the C2 CPU baseline times on the GPU box)."""
import glob
import os

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from enspara_b200 import synth
from oracle import distances as od

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("n,A", [(300, 500), (257, 264), (100, 22), (64, 7), (10, 3)])
def test_sse_rmsd_is_bit_identical_to_the_scalar_lane_version(n, A):
    X = synth.trajectory(n, A, seed=A)
    for ref in (0, n // 2, n - 1):
        a = od.rmsd_f32(X, X[ref])
        b = od.rmsd_f32_sse(X, X[ref])
        assert_array_equal(a, b)
    # and both stay within mdtraj's float32 noise of the float64 truth
    truth = od.rmsd(X, X[0])
    np.testing.assert_allclose(od.rmsd_f32_sse(X, X[0]), truth, rtol=0, atol=5e-3)


@pytest.mark.parametrize("n,A,seed,first", [(500, 500, 0, 0), (333, 264, 7, 12345),
                                            (100, 37, 5, 10 ** 7), (64, 1, 1, 3)])
def test_c_generator_equals_numpy_generator(n, A, seed, first):
    assert_array_equal(od.synth_trajectory(n, A, seed, first),
                       synth.trajectory(n, A, seed=seed, first_frame=first))


def test_c_feature_generator_equals_numpy_generator():
    assert_array_equal(od.synth_features(1000, 64, 3, 7), synth.features(1000, 64, seed=3,
                                                                         first_row=7))
    assert_array_equal(od.synth_features(10, 5, 0, 0), synth.features(10, 5))


def test_use_all_cores_overrides_omp_num_threads(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")       # what torchrun exports
    n = od.use_all_cores()
    assert n == len(os.sched_getaffinity(0)) and od.num_threads() == n


def test_compiled_reference_libdist_loads_standalone():
    from oracle import refharness
    if not glob.glob(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "enspara_geometry",
                                  "libdist*.so")) and not refharness.available():
        pytest.skip("oracle/_ref not built and the reference tree is absent")
    ld = refharness.load_compiled_libdist()
    X = od.synth_features(5000, 64, 1, 0)
    assert_array_equal(ld.euclidean(X, X[17]), od.euclidean(X, X[17]))
    assert_array_equal(ld.manhattan(X, X[17]), od.manhattan(X, X[17]))
