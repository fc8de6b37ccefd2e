"""Host wrappers around the one-vs-all and many-centres kernels.

Everything here enqueues work through the C ABI (enspara_b200/_lib.py) on torch's current
stream and returns host numpy arrays shaped and typed like the reference's.
"""
import numpy as np
import torch

from .. import _lib
from ..device import DeviceFeatures, DeviceTrajectory, ptr, stream_ptr
from ..exception import DataInvalid


def _single_frame_soa(metric, y, like):
    """One datum -> device (SoA coords, trace) for RMSD, or a device row for features."""
    if metric.is_rmsd:
        if isinstance(y, DeviceTrajectory):
            if len(y) != 1:
                raise DataInvalid("expected a single frame, got %d" % len(y))
            return y
        xyz = np.asarray(y.xyz if hasattr(y, "xyz") else y, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[None]
        if xyz.shape[0] != 1:
            xyz = xyz[:1]
        if xyz.shape[1] != like.n_atoms:
            raise DataInvalid("centre has %d atoms, data has %d" % (xyz.shape[1], like.n_atoms))
        return DeviceTrajectory.from_host(xyz)
    arr = np.asarray(y)
    if arr.ndim != 1:
        raise DataInvalid("Target point dimension must be one, got shape %s." % (arr.shape,))
    if arr.shape[0] != like.n_features:
        raise DataInvalid("Target data point dimension (%s) must match data array dimension "
                          "(%s)" % (arr.shape[0], like.n_features))
    if arr.dtype != like.np_dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % (like.np_dtype, arr.dtype))
    return DeviceFeatures.from_host(arr[None])


def one_to_all_device(metric, data, center, exact=True, out=None):
    """Distances of every frame of ``data`` to the single frame ``center`` as a device tensor
    (float32 for RMSD like md.rmsd, float64 for libdist metrics).  No host synchronisation."""
    n = len(data)
    if metric.is_rmsd:
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=data.xyz.device)
        _lib.call("eb_rmsd_one_to_all", ptr(data.xyz), ptr(data.traces), n, data.n_atoms,
                  ptr(center.xyz), ptr(center.traces), ptr(out), int(exact), stream_ptr())
        return out
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=data.X.device)
    from ._engine import _lib_metric
    _lib.call("eb_feat_one_to_all", ptr(data.X), n, data.n_features, data.dt,
              _lib_metric(metric.kind), ptr(center.X), ptr(out), stream_ptr())
    return out


def one_to_all(metric, X, y, out=None):
    """The metric-callable protocol ``d = f(X, y)`` (cluster/util.py:289-313)."""
    data = metric.to_device(X)
    center = _single_frame_soa(metric, y, data)
    d = one_to_all_device(metric, data, center).cpu().numpy()
    if out is not None:
        if out.dtype != np.float64:
            raise DataInvalid("In-place output array must be np.float64, got '%s'." % out.dtype)
        if out.ndim != 1:
            raise DataInvalid("In-place output array must be one-dimensional, got shape %s"
                              % (out.shape,))
        if out.shape[0] != len(data):
            raise DataInvalid("In-place output array dimension (%s) must match number of "
                              "samples in data array (%s)" % (out.shape[0], len(data)))
        out[:] = d
        return out
    return d


def centers_to_device(metric, cluster_centers, like):
    """An iterable of centres (list of 1-frame trajectories / rows, a Trajectory, an array)
    -> one dense device container."""
    if isinstance(cluster_centers, (DeviceTrajectory, DeviceFeatures)):
        return cluster_centers
    if metric.is_rmsd:
        if hasattr(cluster_centers, "xyz"):
            xyz = np.asarray(cluster_centers.xyz, dtype=np.float32)
        else:
            frames = []
            for c in cluster_centers:
                a = np.asarray(c.xyz if hasattr(c, "xyz") else c, dtype=np.float32)
                frames.append(a[0] if a.ndim == 3 else a)
            xyz = np.stack(frames) if frames else np.zeros((0, like.n_atoms, 3), np.float32)
        if xyz.ndim != 3 or xyz.shape[1] != like.n_atoms:
            raise DataInvalid("cluster centres have shape %s, data has %d atoms"
                              % (xyz.shape, like.n_atoms))
        return DeviceTrajectory.from_host(xyz)
    arr = np.asarray([np.asarray(c) for c in cluster_centers]) \
        if not isinstance(cluster_centers, np.ndarray) else cluster_centers
    if arr.ndim != 2 or arr.shape[1] != like.n_features:
        raise DataInvalid("cluster centres have shape %s, data has %d features"
                          % (arr.shape, like.n_features))
    if arr.dtype != like.np_dtype:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % (like.np_dtype, arr.dtype))
    return DeviceFeatures.from_host(arr)


def assign_device(metric, data, centers, frame_idx=None, out_dist=None, out_assign=None,
                  accumulate=False, scatter=False, n_idx=None, k=None, n_dev=None):
    """Nearest-centre pass on the device.  Returns (dist tensor, assign int32 tensor).
    ``scatter``: out arrays are full length and frame f's result goes to position f."""
    n = len(data)
    dev = data.xyz.device if metric.is_rmsd else data.X.device
    m = n if frame_idx is None else int(frame_idx.numel() if n_idx is None else n_idx)
    k = len(centers) if k is None else int(k)
    dt = torch.float32 if metric.is_rmsd else torch.float64
    if out_dist is None:
        out_dist = torch.full((m,), float("inf"), dtype=dt, device=dev)
        out_assign = torch.zeros((m,), dtype=torch.int32, device=dev)
    if m == 0 or k == 0:
        return out_dist, out_assign
    if metric.is_rmsd:
        _lib.call("eb_rmsd_assign_dev", ptr(data.xyz), ptr(data.traces), n, data.n_atoms,
                  ptr(centers.xyz), ptr(centers.traces), k, ptr(frame_idx), m, ptr(out_dist),
                  ptr(out_assign), int(accumulate), int(scatter), ptr(n_dev), stream_ptr())
    else:
        from ._engine import _lib_metric
        _lib.call("eb_feat_assign", ptr(data.X), n, data.n_features, data.dt,
                  _lib_metric(metric.kind), ptr(centers.X), k, ptr(frame_idx), m,
                  ptr(out_dist), ptr(out_assign), int(accumulate), int(scatter), stream_ptr())
    return out_dist, out_assign


#: the tensor-core screen pays off once there are enough centres to amortise the operand split
TC_MIN_CENTERS = 64
#: frames per screen pass (bounds the split-operand scratch: 2 x 12*A_pad bytes per frame)
TC_CHUNK_FRAMES = 262144
#: Error model of the split-FP16 screen: |delta(N * msd)| <= kappa * sqrt(Ga * Gb) with
#: kappa = TC_KAPPA_PER_ATOM * A_pad + TC_KAPPA_CONST.  Derivation (DESIGN.md "K3t error bound"):
#: * entry (i, j) of M is the FP32 sum of 3 * A products (h1 g1, h1 g2, h2 g1 per atom), 16 per
#:   tcgen05.mma.  The tensor core is a block FMA (Fasi, Higham et al. 2021): the 16 products
#:   and the accumulator are aligned to the largest exponent, each addend truncated to 24 + e
#:   bits, summed exactly, and the result truncated to FP32.  e = TC_ALIGN_EXTRA_BITS = 2 on
#:   B200 -- measured by tests/test_gpu_tc_screen.py::test_tensor_core_keeps_two_alignment_bits
#:   (products of 1/2 and 1/4 ulp of the largest addend survive, 1/8 ulp does not), so the
#:   claim is re-checked on the device the tests run on.  Every magnitude involved is bounded
#:   by B_ij = sum_a |x_ia| |y_ja|, so one MMA errs by at most (16 * 2^-e + 1) u B_ij with
#:   u = 2^-23, and the 3 A / 16 MMAs of an entry by 3 A (16 * 2^-e + 1) / 16 u B_ij
#:   = 0.9375 A u B_ij;
#: * representation: |x - (h1 + h2) / 2^8| <= 2^-24 |x| (+ 2^-33 nm absolute once h2 is
#:   subnormal), dropped h2 g2 <= 2^-22 |x| |y|: together < 3 u B_ij (the constant term);
#: * Cauchy-Schwarz: B_ij <= sqrt(Gx_i Gy_j), hence ||dM||_F <= (0.9375 A + 3) u sqrt(Ga Gb);
#: * lambda_max = max over rotations R of tr(R^T M) (Horn), so |d lambda| <= ||dM||_* <=
#:   sqrt(3) ||dM||_F, and N * msd = Ga + Gb - 2 lambda doubles it.
#: Measured worst case on B200 (tests/test_gpu_tc_screen.py::test_tc_error_bound, natural and
#: adversarial same-sign inputs): about 1/4 of this bound.  Round 1 used 8 * 2^-24 per atom,
#: 8x a measurement on natural data but not a bound.
TC_ALIGN_EXTRA_BITS = 2
TC_ADDENDS_PER_ATOM = 3.0 * (16.0 * 2.0 ** -TC_ALIGN_EXTRA_BITS + 1.0) / 16.0      # 0.9375
TC_KAPPA_PER_ATOM = 2.0 * 3.0 ** 0.5 * TC_ADDENDS_PER_ATOM * 2.0 ** -23
TC_KAPPA_CONST = 2.0 * 3.0 ** 0.5 * 3.0 * 2.0 ** -23


def tc_kappa(a_pad):
    return TC_KAPPA_PER_ATOM * a_pad + TC_KAPPA_CONST


#: Built-in audit of the screen (ENSPARA_B200_TC_AUDIT: 0 off, 1 default, 2 every call): a random
#: sample of the frames of a pass is scored against ALL centres with the exact float64 kernel
#: and must reproduce the distance and the assignment the screen + re-score produced, bit for
#: bit; a mismatch raises.  Dense passes (>= TC_AUDIT_MIN_PAIRS frame-centre pairs) are always
#: audited; small passes (PAM's per-proposal subsets) once every TC_AUDIT_EVERY calls.
TC_AUDIT_FRAMES = 64
TC_AUDIT_MIN_PAIRS = 1 << 24
TC_AUDIT_EVERY = 64
TC_AUDIT_LIST = 256
_audit_calls = [0]
audit_stats = {"passes_audited": 0, "frames_audited": 0}


def _audit_level():
    import os
    return int(os.environ.get("ENSPARA_B200_TC_AUDIT", "1"))


def _audit_tc(data, centers, k, m, frame_idx, scatter, out_dist, out_assign, ws):
    """Exact-score a sample of the pass's frames against every centre and compare."""
    dev = data.xyz.device
    S = min(TC_AUDIT_FRAMES, m)
    L = TC_AUDIT_LIST
    chunks = (k + L - 1) // L
    key = ("audit", k, S)
    if ws.get("audit_key") != key:
        ch = torch.arange(chunks, device=dev, dtype=torch.int32)
        lists = (ch[:, None] * L + torch.arange(L, device=dev, dtype=torch.int32)[None, :])
        ws["audit_lists"] = lists.repeat(S, 1).contiguous()                    # (S*chunks, L)
        ws["audit_count"] = torch.clamp(k - ch * L, max=L).to(torch.int32).repeat(S).contiguous()
        ws["audit_lo"] = torch.full((S * chunks * L,), float("-inf"), dtype=torch.float32,
                                    device=dev)
        ws["audit_up"] = torch.full((S * chunks,), float("inf"), dtype=torch.float32, device=dev)
        ws["audit_zero"] = torch.zeros(4, dtype=torch.int32, device=dev)
        ws["audit_d"] = torch.empty(S * chunks, dtype=torch.float32, device=dev)
        ws["audit_a"] = torch.empty(S * chunks, dtype=torch.int32, device=dev)
        ws["audit_flag"] = torch.empty(S * chunks, dtype=torch.int32, device=dev)
        ws["audit_gen"] = torch.Generator(device=dev)
        ws["audit_gen"].manual_seed(0x5EED)
        ws["audit_key"] = key
    pos = torch.randint(0, m, (S,), device=dev, generator=ws["audit_gen"])
    src = pos if frame_idx is None else frame_idx[:m][pos]
    res = src if (scatter and frame_idx is not None) else pos
    fidx = src.to(torch.int64).repeat_interleave(chunks).contiguous()
    _lib.call("eb_rmsd_score_lists", ptr(data.xyz), ptr(data.traces), S * chunks, data.n_atoms,
              ptr(centers.xyz), ptr(centers.traces), ptr(ws["audit_count"]),
              ptr(ws["audit_lists"]), L, ptr(ws["audit_lo"]), ptr(ws["audit_up"]),
              ptr(ws["audit_zero"]), ptr(fidx), ptr(ws["audit_d"]), ptr(ws["audit_a"]),
              ptr(ws["audit_flag"]), stream_ptr())
    # nearest over the chunks, lowest centre index on ties: (distance bits, index) as one key
    keyv = (ws["audit_d"].view(torch.int32).to(torch.int64) << 32) | ws["audit_a"].to(torch.int64)
    best = keyv.view(S, chunks).min(dim=1).values
    want_d = (best >> 32).to(torch.int32).view(torch.float32)
    want_a = (best & 0xFFFFFFFF).to(torch.int32)
    bad = (want_d != out_dist[res]) | (want_a != out_assign[res])
    audit_stats["passes_audited"] += 1
    audit_stats["frames_audited"] += S
    if bool(bad.any().item()):
        i = int(torch.nonzero(bad)[0])
        raise RuntimeError(
            "enspara_b200: tensor-core screen audit FAILED for frame %d: screen path gave "
            "(centre %d, %.9g nm), exact path (centre %d, %.9g nm).  Set "
            "ENSPARA_B200_TC_MIN_CENTERS to a huge value to force the exact kernel and please "
            "report this." % (int(src[i]), int(out_assign[res][i]), float(out_dist[res][i]),
                              int(want_a[i]), float(want_d[i])))


def tc_applicable(metric, data, k):
    import os
    min_k = int(os.environ.get("ENSPARA_B200_TC_MIN_CENTERS", TC_MIN_CENTERS))
    return metric.is_rmsd and k >= min_k and len(data) >= 1


def assign_device_tc(metric, data, centers, k=None, stats=None, frame_idx=None, n_idx=None,
                     out_dist=None, out_assign=None, scatter=False, workspace=None,
                     n_dev=None, defer=False):
    """Nearest-centre pass through the tcgen05 screen + exact re-score (csrc/eb_tc_screen.cu).
    Same result as ``assign_device`` (the exact path decides).  ``frame_idx`` (device int64)
    restricts the pass to a subset; with ``scatter`` the full-length ``out_*`` arrays receive
    frame f's result at position f (PAM's re-assignment of X[dst_up_assig_this],
    kmedoids.py:666-667).  ``workspace``: dict reused between calls (scratch buffers).

    ``n_dev`` (device int32/int64 tensor, first element read as int32): the real subset size
    when only the device knows it; ``n_idx`` is then the host's upper bound.  ``defer=True``
    (single-chunk subsets only) returns without synchronising: the third return value is a
    device int32 counter of frames whose candidate lists overflowed -- the caller reads it with
    its next read-back and, if non-zero, calls ``assign_device`` for the subset (the exact
    fallback this function otherwise does itself)."""
    n = len(data)
    k = len(centers) if k is None else int(k)
    dev = data.xyz.device
    lib = _lib.load()
    m = n if frame_idx is None else int(frame_idx.numel() if n_idx is None else n_idx)
    if out_dist is None:
        out_dist = torch.full((m,), float("inf"), dtype=torch.float32, device=dev)
        out_assign = torch.zeros((m,), dtype=torch.int32, device=dev)
        scatter = False
    if m == 0 or k == 0:
        return (out_dist, out_assign, None) if defer else (out_dist, out_assign)
    ws = workspace if workspace is not None else {}
    chunk = min(m, TC_CHUNK_FRAMES)
    need = int(lib.eb_tc_scratch_bytes(chunk, data.n_atoms, k))
    if ws.get("scratch") is None or ws["scratch"].numel() < need:
        ws["scratch"] = None
        ws["scratch"] = torch.empty(need + need // 4, dtype=torch.uint8, device=dev)
    if ws.get("cand") is None or ws["cand"].numel() < m:
        ws["cand"] = torch.empty(m + m // 4 + 128, dtype=torch.int32, device=dev)
    scratch, cand = ws["scratch"], ws["cand"][:m]
    kappa = tc_kappa(data.a_pad)
    if defer:
        if frame_idx is None or m > chunk:
            raise ValueError("defer=True needs a single-chunk frame subset")
        if ws.get("ovf") is None:
            ws["ovf"] = torch.zeros(1, dtype=torch.int32, device=dev)
        ws["ovf"].zero_()
        od = out_dist
        oa = out_assign
        _lib.call("eb_rmsd_assign_tc_dev", ptr(data.xyz), ptr(data.traces), m, data.n_atoms,
                  ptr(centers.xyz), ptr(centers.traces), k, float(kappa), ptr(frame_idx),
                  int(scatter), ptr(od), ptr(oa), ptr(cand), ptr(scratch), None, 1, ptr(n_dev),
                  ptr(ws["ovf"]), stream_ptr())
        return out_dist, out_assign, ws["ovf"]
    for lo in range(0, m, chunk):
        hi = min(m, lo + chunk)
        if frame_idx is None:
            _lib.call("eb_rmsd_assign_tc", ptr(data.xyz[lo:]), ptr(data.traces[lo:]), hi - lo,
                      data.n_atoms, ptr(centers.xyz), ptr(centers.traces), k, float(kappa),
                      None, 0, ptr(out_dist[lo:]), ptr(out_assign[lo:]), ptr(cand[lo:]),
                      ptr(scratch), None, 1, stream_ptr())
        else:
            od = out_dist if scatter else out_dist[lo:]
            oa = out_assign if scatter else out_assign[lo:]
            _lib.call("eb_rmsd_assign_tc", ptr(data.xyz), ptr(data.traces), hi - lo,
                      data.n_atoms, ptr(centers.xyz), ptr(centers.traces), k, float(kappa),
                      ptr(frame_idx[lo:]), int(scatter), ptr(od), ptr(oa), ptr(cand[lo:]),
                      ptr(scratch), None, 1, stream_ptr())
    overflow = torch.nonzero(cand < 0).view(-1)          # positions within the pass
    if stats is not None:
        stats["survivors_mean"] = float(cand.clamp(min=0).float().mean().cpu())
        stats["overflow_frames"] = int(overflow.numel())
    if overflow.numel() > 0:
        # a candidate list overflowed for these frames: exact pass over all centres
        if frame_idx is None:
            assign_device(metric, data, centers, frame_idx=overflow.to(torch.int64).contiguous(),
                          out_dist=out_dist, out_assign=out_assign, accumulate=False,
                          scatter=True, k=k)
        elif scatter:
            assign_device(metric, data, centers,
                          frame_idx=frame_idx[:m][overflow].contiguous(), out_dist=out_dist,
                          out_assign=out_assign, accumulate=False, scatter=True, k=k)
        else:
            d2, a2 = assign_device(metric, data, centers,
                                   frame_idx=frame_idx[:m][overflow].contiguous(), k=k)
            out_dist[overflow] = d2
            out_assign[overflow] = a2
    level = _audit_level()
    if level > 0:
        _audit_calls[0] += 1
        if (level >= 2 or m * k >= TC_AUDIT_MIN_PAIRS
                or _audit_calls[0] % TC_AUDIT_EVERY == 0):
            _audit_tc(data, centers, k, m, frame_idx, scatter, out_dist, out_assign, ws)
    return out_dist, out_assign


def assign_device_auto(metric, data, centers, k=None):
    kk = len(centers) if k is None else int(k)
    if tc_applicable(metric, data, kk):
        return assign_device_tc(metric, data, centers, k=kk)
    return assign_device(metric, data, centers, k=kk)


def assign_host(metric, trajectory, cluster_centers):
    """assign_to_nearest_center (cluster/util.py:159-205) -> (int64[n], float64[n]) numpy."""
    data = metric.to_device(trajectory)
    centers = centers_to_device(metric, cluster_centers, data)
    d, a = assign_device_auto(metric, data, centers)
    return a.cpu().numpy().astype(np.int64), d.cpu().numpy().astype(np.float64)
