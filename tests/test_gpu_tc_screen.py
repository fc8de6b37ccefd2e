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
    """Many identical centres -> a candidate list (64 entries; a list covers 120 centres here)
    overflows for every frame -> exact fallback, same result."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    data = DeviceTrajectory.from_host(synth.trajectory(9600, 64, seed=2))
    cen = data.gather(np.array([5] * 400 + [17] * 400, dtype=np.int64))
    d0, a0 = _ops.assign_device(util.RMSD, data, cen)
    stats = {}
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen, stats=stats)
    assert stats["overflow_frames"] > 0
    assert cuda.equal(a0, a1) and cuda.equal(d0, d1)
    assert set(np.unique(a1.cpu().numpy())) <= {0, 400}     # lowest index among duplicates


@pytest.mark.parametrize("n,A,k,m", [(6000, 500, 200, 700), (3000, 48, 1000, 37),
                                     (2000, 240, 64, 2000), (5000, 96, 333, 129)])
def test_tc_subset_scatter_equals_exact(cuda, n, A, k, m):
    """Frame subsets (PAM's X[dst_up_assig_this], kmedoids.py:666-667): the screen restricted to
    `frame_idx` with scattered outputs equals the exact kernel on the same subset, and leaves
    every other frame's entry untouched."""
    torch = cuda
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    data = DeviceTrajectory.from_host(synth.trajectory(n, A, seed=n + k))
    cen = data.gather(np.linspace(0, n - 1, k).astype(np.int64))
    rs = np.random.RandomState(m)
    idx = torch.as_tensor(np.sort(rs.choice(n, m, replace=False)), device="cuda")
    buf = torch.cat([idx, torch.zeros(50, dtype=torch.int64, device="cuda")])  # longer buffer
    d0 = torch.full((n,), -7.0, device="cuda")
    a0 = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    d1, a1 = d0.clone(), a0.clone()
    _ops.assign_device(util.RMSD, data, cen, frame_idx=buf, n_idx=m, out_dist=d0, out_assign=a0,
                       scatter=True)
    ws = {}
    for _ in range(2):       # second call reuses the workspace
        _ops.assign_device_tc(util.RMSD, data, cen, frame_idx=buf, n_idx=m, out_dist=d1,
                              out_assign=a1, scatter=True, workspace=ws)
    assert torch.equal(a0, a1) and torch.equal(d0, d1)
    untouched = torch.ones(n, dtype=torch.bool, device="cuda")
    untouched[idx] = False
    assert bool((d1[untouched] == -7.0).all()) and bool((a1[idx] >= 0).all())
    # compact (non-scatter) form
    d2, a2 = _ops.assign_device_tc(util.RMSD, data, cen, frame_idx=idx)
    assert torch.equal(d2, d1[idx]) and torch.equal(a2, a1[idx])


def test_pam_sweep_uses_tc_and_matches_exact_path(cuda):
    """A PAM sweep with enough medoids for the tensor-core subset path gives exactly what the
    exact-kernel sweep gives (medoids, assignments, distances) and what the oracle gives."""
    from enspara_b200 import synth
    from enspara_b200.cluster import kcenters, util
    from enspara_b200.cluster._pam import PamEngine
    from enspara_b200.cluster.kcenters import _SingleComm
    from enspara_b200.device import DeviceTrajectory
    from oracle import cluster as oc
    from oracle import distances as od
    X = synth.trajectory(4000, 32, seed=21)
    data = DeviceTrajectory.from_host(X)
    r = kcenters.kcenters(data, "rmsd", n_clusters=80)
    ctr = [int(c) for c in r.center_indices]
    out = []
    for use_tc in (True, False):
        pam = PamEngine(data, util.RMSD, _SingleComm(), r.distances, r.assignments, ctr)
        assert pam.use_tc
        pam.use_tc = use_tc
        pam.TC_MIN_PAIRS = 1
        acc = pam.sweep(random_state=3)
        a, d = pam.results_host()
        out.append((list(pam.medoid_global), a, d, acc))
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3]
    assert_array_equal(out[0][1], out[1][1])
    assert_array_equal(out[0][2], out[1][2])
    T = od.Trajectory(X)
    ind, d, a, _ = oc.pam_update(T, od.rmsd, list(ctr), r.assignments.copy(),
                                 r.distances.copy(), random_state=3)
    assert [int(i) for i in ind] == out[0][0]
    assert_array_equal(a, out[0][1])
    assert_allclose(d, out[0][2], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n,A,k", [(3000, 22, 100), (2500, 264, 70), (1000, 13, 64), (50, 30, 200)])
def test_tc_any_atom_count(cuda, n, A, k):
    """Atom counts whose padded row is not a multiple of 16 (22 -> 24, 264, 13 -> 16): the
    packed operand images are zero-padded, the result still equals the exact kernel's."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    data = DeviceTrajectory.from_host(synth.trajectory(n, A, seed=A))
    cen = data.gather(np.linspace(0, n - 1, k).astype(np.int64))
    assert _ops.tc_applicable(util.RMSD, data, k)
    d0, a0 = _ops.assign_device(util.RMSD, data, cen)
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen)
    assert cuda.equal(a0, a1) and cuda.equal(d0, d1)
