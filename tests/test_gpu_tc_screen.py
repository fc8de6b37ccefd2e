"""GPU parity: the tcgen05 3xTF32 screen + exact re-score returns EXACTLY what the exact float64
many-centres kernel returns (assignments and float32 distances, bit for bit), and the oracle's
result within the RMSD tolerance."""
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


@pytest.mark.parametrize("n,A,k", [(5000, 500, 200), (3000, 240, 64), (1000, 48, 333),
                                   (129, 500, 65), (4096, 1000, 96)])
def test_tc_assign_equals_exact_assign(cuda, n, A, k):
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    data = DeviceTrajectory.from_host(synth.trajectory(n, A, seed=A + k))
    cen = data.gather(np.linspace(0, n - 1, k).astype(np.int64))
    assert _ops.tc_applicable(util.RMSD, data, k)
    d0, a0 = _ops.assign_device(util.RMSD, data, cen)
    stats = {}
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen, stats=stats)
    assert cuda.equal(a0, a1)
    assert cuda.equal(d0, d1)
    assert stats["survivors_mean"] >= 1.0


def test_tc_assign_matches_oracle(cuda):
    from enspara_b200 import synth
    from enspara_b200.cluster import util
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.trajectory(1500, 96, seed=5)
    T = od.Trajectory(X)
    idx = np.arange(0, 1500, 15)          # 100 centres -> tensor-core path
    want_a, want_d = oc.assign_to_nearest_center(T, [T[i] for i in idx], od.rmsd)
    got_a, got_d = util.assign_to_nearest_center(T, [T[i] for i in idx], "rmsd")
    assert_array_equal(got_a, want_a)
    assert_allclose(got_d, want_d, rtol=1e-5, atol=1e-6)


def test_tc_overflow_falls_back_to_exact(cuda):
    """Many identical centres -> every frame has > 128 survivors -> exact fallback, same result."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    data = DeviceTrajectory.from_host(synth.trajectory(600, 64, seed=2))
    cen = data.gather(np.array([5] * 150 + [17] * 150, dtype=np.int64))
    d0, a0 = _ops.assign_device(util.RMSD, data, cen)
    stats = {}
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen, stats=stats)
    assert stats["overflow_frames"] > 0
    assert cuda.equal(a0, a1) and cuda.equal(d0, d1)
    assert set(np.unique(a1.cpu().numpy())) <= {0, 150}     # lowest index among duplicates
