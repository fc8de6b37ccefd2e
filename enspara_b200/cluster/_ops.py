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
                  accumulate=False, scatter=False, n_idx=None, k=None):
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
        _lib.call("eb_rmsd_assign", ptr(data.xyz), ptr(data.traces), n, data.n_atoms,
                  ptr(centers.xyz), ptr(centers.traces), k, ptr(frame_idx), m, ptr(out_dist),
                  ptr(out_assign), int(accumulate), int(scatter), stream_ptr())
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
#: error model of the 3xTF32 screen: |N * d(msd)| <= TC_KAPPA_PER_ATOM * A_pad * sqrt(Ga*Gb).
#: Measured on B200 (scripts/dev_tc_debug.py): the worst pair is 1.0 * 2^-24 per atom (FP32
#: accumulation in the tensor core truncates); 8x margin.
TC_KAPPA_PER_ATOM = 8.0 * 2.0 ** -24


def tc_applicable(metric, data, k):
    return metric.is_rmsd and k >= TC_MIN_CENTERS and len(data) >= 1


def assign_device_tc(metric, data, centers, k=None, stats=None, frame_idx=None, n_idx=None,
                     out_dist=None, out_assign=None, scatter=False, workspace=None):
    """Nearest-centre pass through the tcgen05 screen + exact re-score (csrc/eb_tc_screen.cu).
    Same result as ``assign_device`` (the exact path decides).  ``frame_idx`` (device int64)
    restricts the pass to a subset; with ``scatter`` the full-length ``out_*`` arrays receive
    frame f's result at position f (PAM's re-assignment of X[dst_up_assig_this],
    kmedoids.py:666-667).  ``workspace``: dict reused between calls (scratch buffers)."""
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
        return out_dist, out_assign
    ws = workspace if workspace is not None else {}
    chunk = min(m, TC_CHUNK_FRAMES)
    need = int(lib.eb_tc_scratch_bytes(chunk, data.n_atoms, k))
    if ws.get("scratch") is None or ws["scratch"].numel() < need:
        ws["scratch"] = None
        ws["scratch"] = torch.empty(need + need // 4, dtype=torch.uint8, device=dev)
    if ws.get("cand") is None or ws["cand"].numel() < m:
        ws["cand"] = torch.empty(m + m // 4 + 128, dtype=torch.int32, device=dev)
    scratch, cand = ws["scratch"], ws["cand"][:m]
    kappa = TC_KAPPA_PER_ATOM * data.a_pad
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
