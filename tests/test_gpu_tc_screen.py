"""GPU parity: the tcgen05 split-FP16 screen + exact re-score returns EXACTLY what the exact
float64 many-centres kernel returns (assignments and float32 distances, bit for bit), and the
ORACLE's result (restated mdtraj RMSD driven by the reference's assign loop,
cluster/util.py:159-205) within the RMSD tolerance -- including at the BASELINE config-5 shape
(500 atoms, >= 1000 centres).  Also: the error bound the screen's discard rule rests on
(cluster/_ops.py TC_KAPPA) holds entry by entry, and the built-in audit runs."""
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


def _oracle_assign_sample(X, centre_idx, sample, rtol=1e-5, atol=1e-6):
    """GPU assign of ALL frames of X to the centres X[centre_idx] (tensor-core path) vs the
    oracle on the sampled frames.  Index mismatches are allowed only as documented near-ties
    (< 1e-6 nm between the two candidates, north_star)."""
    from enspara_b200.cluster import util
    from oracle import cluster as oc
    from oracle import distances as od
    T = od.Trajectory(X)
    centers = [T[int(i)] for i in centre_idx]
    got_a, got_d = util.assign_to_nearest_center(T, centers, "rmsd")
    S = od.Trajectory(X[sample])
    want_a, want_d = oc.assign_to_nearest_center(S, centers, od.rmsd)
    assert_allclose(got_d[sample], want_d, rtol=rtol, atol=atol)
    diff = np.nonzero(got_a[sample] != want_a)[0]
    for i in diff:      # a different index must be a tie below 1e-6 nm
        other = od.rmsd(od.Trajectory(X[sample[i]][None]), centers[int(got_a[sample][i])])[0]
        assert abs(float(other) - float(want_d[i])) < 1e-6
    return len(diff)


def test_tc_assign_matches_oracle_at_config5_shape(cuda):
    """BASELINE configs[4] shape: 500 atoms, 1200 centres (25 centre tiles, several segments);
    60 000 frames through the screen, 300 of them checked against the oracle."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops
    n, A, k = 60_000, 500, 1200
    X = synth.device_trajectory_aos(n, A, seed=0).cpu().numpy()
    rs = np.random.RandomState(1)
    centre_idx = np.sort(rs.choice(n, k, replace=False))
    sample = np.sort(rs.choice(n, 300, replace=False))
    before = dict(_ops.audit_stats)
    _oracle_assign_sample(X, centre_idx, sample)
    # the dense pass was audited (>= 2^24 pairs): 64 frames exact-scored against all centres
    assert _ops.audit_stats["passes_audited"] > before["passes_audited"]


@pytest.mark.parametrize("n,A,k", [(20_000, 264, 1000), (30_000, 22, 2000)])
def test_tc_assign_matches_oracle_other_atom_counts(cuda, n, A, k):
    """Config-1-like (264 atoms) and frame0-like (22 atoms) molecules with >= 1000 centres."""
    from enspara_b200 import synth
    X = synth.device_trajectory_aos(n, A, seed=3).cpu().numpy()
    rs = np.random.RandomState(2)
    centre_idx = np.sort(rs.choice(n, k, replace=False))
    sample = np.sort(rs.choice(n, 400, replace=False))
    _oracle_assign_sample(X, centre_idx, sample)


def test_tc_error_bound(cuda):
    """The inequality the discard rule is derived from, checked entry by entry on the
    accumulators the tensor core really produced (mode 0 dumps them):
        |M_tc[f,c,i,j] - M_exact[f,c,i,j]| <= (0.9375 A + 3) 2^-23 sqrt(Gx_i(f) Gy_j(c)),
    and its consequence for N * msd with kappa = _ops.tc_kappa(A_pad).  Reports how far the
    observed worst case is from the bound (expected: well below 1)."""
    torch = cuda
    from enspara_b200 import _lib, synth
    from enspara_b200.cluster import _ops
    from enspara_b200.device import DeviceTrajectory, ptr, stream_ptr
    L = _lib.load()
    worst_entry = worst_msd = 0.0
    for n, k, A, seed in ((1024, 96, 500, 3), (512, 144, 1000, 4), (2048, 48, 264, 5),
                          (700, 50, 22, 6), (256, 48, 500, -1), (256, 48, 1000, -2)):
        if seed >= 0:
            data = DeviceTrajectory.from_host(synth.trajectory(n, A, seed=seed))
            cen = data.gather(np.linspace(0, n - 1, k).astype(np.int64))
        else:
            # adversarial for a truncating accumulator: every product positive, mantissas
            # full of ones, magnitudes alike -- the running sum grows monotonically and every
            # addition truncates in the same direction (uncentred on purpose: mode 0 only
            # multiplies)
            g = torch.Generator(device="cuda")
            g.manual_seed(-seed)
            a_pad = int(L.eb_rmsd_apad(A))

            def rows(m):
                v = 1.0 + torch.rand((m, 3, a_pad), device="cuda", generator=g)
                v = torch.nextafter(v.float(), torch.zeros_like(v).float())
                v[:, :, A:] = 0
                return v.contiguous()
            fx, cx = rows(n), rows(k)
            data = DeviceTrajectory(fx, (fx.double() ** 2).sum(dim=(1, 2)), A)
            cen = DeviceTrajectory(cx, (cx.double() ** 2).sum(dim=(1, 2)), A)
        scratch = torch.zeros(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8,
                              device="cuda")
        dbg = torch.zeros((n, k, 9), dtype=torch.float32, device="cuda")
        _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
                  ptr(cen.traces), k, 0.0, None, 0, None, None, None, ptr(scratch), ptr(dbg), 0,
                  stream_ptr())
        xs = data.xyz.double()                               # (n, 3, A_pad) centred
        cs = cen.xyz.double()
        M = torch.einsum("fia,cja->fcij", xs, cs)
        dM = (dbg.double().view(n, k, 3, 3) - M).abs()
        gx = (xs * xs).sum(dim=2)                            # (n, 3) per-coordinate traces
        gy = (cs * cs).sum(dim=2)
        bound = (_ops.TC_ADDENDS_PER_ATOM * data.a_pad + 3.0) * 2.0 ** -23 * torch.sqrt(
            gx[:, None, :, None] * gy[None, :, None, :])
        ratio = float((dM / bound.clamp(min=1e-300)).max())
        worst_entry = max(worst_entry, ratio)
        assert ratio < 1.0, "entry bound violated (n=%d k=%d A=%d): %.3g" % (n, k, A, ratio)
        # consequence: |d lambda| <= sqrt(3) ||dM||_F  ->  |d(N msd)| <= kappa sqrt(Ga Gb)
        fro = torch.sqrt((dM * dM).sum(dim=(2, 3)))
        lhs = 2.0 * 3.0 ** 0.5 * fro
        rhs = _ops.tc_kappa(data.a_pad) * torch.sqrt(data.traces[:, None] * cen.traces[None, :])
        worst_msd = max(worst_msd, float((lhs / rhs).max()))
        assert bool((lhs <= rhs).all())
    print("tc error bound: worst entry ratio %.4f, worst N*msd ratio %.4f of the bound"
          % (worst_entry, worst_msd))
    assert worst_entry < 0.6        # the bound is meant to have margin, not to be tight


def test_tensor_core_keeps_two_alignment_bits(cuda):
    """The hardware property TC_KAPPA rests on (cluster/_ops.py TC_ALIGN_EXTRA_BITS = 2): when
    tcgen05.mma.kind::f16 aligns the 16 products of a k-step (and the accumulator) to the
    largest exponent, an addend keeps at least two bits below the FP32 ulp of the largest one.
    One product of 2^30 (ulp 2^7) plus fifteen products of 2^(6-t): with >= t+1 extra bits they
    survive the alignment and show up (truncated to the ulp) in the result."""
    torch = cuda
    from enspara_b200 import _lib
    from enspara_b200.cluster import _ops
    from enspara_b200.device import DeviceTrajectory, ptr, stream_ptr
    L = _lib.load()
    A, n, k = 32, 128, 48

    def probe(t, other_step):
        fx = torch.zeros((n, 3, A), dtype=torch.float32, device="cuda")
        cx = torch.zeros((k, 3, A), dtype=torch.float32, device="cuda")
        fx[:, 0, 0] = 2.0 ** 7                    # x * 2^8 = 2^15 -> product 2^30
        cx[:, 0, 0] = 2.0 ** 7
        lo = 16 if other_step else 1              # the other / the same 16-atom MMA k-step
        fx[:, 0, lo:lo + 15] = 2.0 ** -5
        cx[:, 0, lo:lo + 15] = 2.0 ** (-5 - t)    # fifteen products of 2^(6-t)
        one_n = torch.ones(n, dtype=torch.float64, device="cuda")
        data = DeviceTrajectory(fx, one_n, A)
        cen = DeviceTrajectory(cx, one_n[:k].clone(), A)
        scratch = torch.zeros(int(L.eb_tc_scratch_bytes(n, A, k)), dtype=torch.uint8,
                              device="cuda")
        dbg = torch.zeros((n, k, 9), dtype=torch.float32, device="cuda")
        _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), n, A, ptr(cen.xyz),
                  ptr(cen.traces), k, 0.0, None, 0, None, None, None, ptr(scratch), ptr(dbg), 0,
                  stream_ptr())
        got = dbg[:, :, 0].double() * 65536.0 - 2.0 ** 30
        assert bool((got == got[0, 0]).all())     # every lane / column behaves alike
        return float(got[0, 0])
    assert _ops.TC_ALIGN_EXTRA_BITS == 2
    for other in (False, True):
        assert probe(0, other) == 896.0           # 15 * 64 = 960 -> 7 ulps survive
        assert probe(1, other) == 384.0           # 15 * 32 = 480 -> 3 ulps: two extra bits kept
        assert probe(2, other) in (0.0, 128.0)    # a third extra bit would be a bonus, not needed


def test_tc_audit_every_call_and_detects_corruption(cuda, monkeypatch):
    """ENSPARA_B200_TC_AUDIT=2 audits every pass, subsets included; an assignment the exact
    path disagrees with raises."""
    torch = cuda
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    monkeypatch.setenv("ENSPARA_B200_TC_AUDIT", "2")
    data = DeviceTrajectory.from_host(synth.trajectory(4000, 96, seed=8))
    cen = data.gather(np.linspace(0, 3999, 300).astype(np.int64))
    before = _ops.audit_stats["passes_audited"]
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen)
    idx = torch.arange(5, 3000, 7, device="cuda")
    d = torch.full((4000,), -1.0, device="cuda")
    a = torch.full((4000,), -1, dtype=torch.int32, device="cuda")
    _ops.assign_device_tc(util.RMSD, data, cen, frame_idx=idx, out_dist=d, out_assign=a,
                          scatter=True)
    assert _ops.audit_stats["passes_audited"] == before + 2
    assert torch.equal(a[idx], a1[idx]) and torch.equal(d[idx], d1[idx])
    # corrupt what the audit compares against: it must notice
    ws = {}
    real = _ops._audit_tc

    def corrupting(data_, centers_, k_, m_, fi_, sc_, od_, oa_, ws_):
        oa_.add_(1)
        return real(data_, centers_, k_, m_, fi_, sc_, od_, oa_, ws_)
    monkeypatch.setattr(_ops, "_audit_tc", corrupting)
    with pytest.raises(RuntimeError, match="audit FAILED"):
        _ops.assign_device_tc(util.RMSD, data, cen, workspace=ws)


def test_tc_fp16_range_overflow_falls_back_to_exact(cuda):
    """Coordinates beyond the FP16 split's range (|x| * 2^8 > 65000, i.e. > 253 nm from the
    centroid) cannot go through the screen: the pack kernel flags it and every frame takes the
    exact kernel -- same result."""
    from enspara_b200 import synth
    from enspara_b200.cluster import _ops, util
    from enspara_b200.device import DeviceTrajectory
    X = synth.trajectory(2000, 40, seed=9) * np.float32(200.0)       # extents of several hundred nm
    assert np.abs(X - X.mean(axis=1, keepdims=True)).max() > 260.0
    data = DeviceTrajectory.from_host(X)
    cen = data.gather(np.linspace(0, 1999, 70).astype(np.int64))
    d0, a0 = _ops.assign_device(util.RMSD, data, cen)
    stats = {}
    d1, a1 = _ops.assign_device_tc(util.RMSD, data, cen, stats=stats)
    assert stats["overflow_frames"] == 2000
    assert cuda.equal(a0, a1) and cuda.equal(d0, d1)
