"""Device-resident containers: how frames live in HBM.

``DeviceTrajectory``  frames as PRE-CENTRED float32 SoA blocks (n, 3, A_pad) + float64 traces.
                      Built once from an ``md.Trajectory``-like object (anything with ``.xyz``)
                      or an (n, A, 3) array; replaces the copy + centre + trace that
                      ``mdtraj.rmsd`` repeats on every call.
``DeviceFeatures``    an (n, F) row-major matrix in its own dtype (libdist's fused types,
                      /root/reference/enspara/geometry/libdist.pyx:9-15).

PyTorch is used for device memory, streams and pinned staging only; all arithmetic is in
libenspara_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .exception import DataInvalid


def cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "enspara_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor as c_void_p (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


_NP_TO_DT = {np.dtype(np.float32): _lib.DT_F32, np.dtype(np.float64): _lib.DT_F64,
             np.dtype(np.int8): _lib.DT_I8, np.dtype(np.int16): _lib.DT_I16,
             np.dtype(np.int32): _lib.DT_I32, np.dtype(np.int64): _lib.DT_I64}
_NP_TO_TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16,
                np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}


def _h2d(host_np, dev_tensor):
    """Copy a contiguous numpy array into a device tensor of the same size (via pinned
    staging when the array is large, plain copy otherwise)."""
    src = torch.from_numpy(host_np)
    dev_tensor.view(-1).copy_(src.view(-1), non_blocking=False)


class DeviceTrajectory:
    """Pre-centred SoA frames + traces in HBM."""

    #: frames converted per staging chunk when uploading from the host
    CHUNK_BYTES = 1 << 28

    def __init__(self, xyz_soa, traces, n_atoms, topology=None, host=None):
        self.xyz = xyz_soa          # float32 (n, 3, A_pad)
        self.traces = traces        # float64 (n,)
        self.n_atoms = int(n_atoms)
        self.a_pad = int(xyz_soa.shape[2])
        self.top = self.topology = topology
        self.host = host            # the caller's object, used to hand back centres

    def __len__(self):
        return int(self.xyz.shape[0])

    @property
    def frame_bytes(self):
        return 12 * self.a_pad

    @classmethod
    def empty(cls, n, n_atoms, topology=None):
        dev = cuda_device()
        a_pad = _lib.load().eb_rmsd_apad(int(n_atoms))
        xyz = torch.empty((n, 3, a_pad), dtype=torch.float32, device=dev)
        tr = torch.empty((n,), dtype=torch.float64, device=dev)
        return cls(xyz, tr, n_atoms, topology)

    @classmethod
    def from_host(cls, traj, precentered=False):
        """Upload an ``md.Trajectory``-like object or an (n, A, 3) float array."""
        xyz = traj.xyz if hasattr(traj, "xyz") else traj
        xyz = np.asarray(xyz)
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise DataInvalid(
                "RMSD clustering needs coordinates of shape (n_frames, n_atoms, 3), got %s."
                % (xyz.shape,))
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n, A = xyz.shape[0], xyz.shape[1]
        out = cls.empty(n, A, getattr(traj, "top", None))
        out.host = traj
        if n == 0:
            return out
        # double-buffered upload: the H2D copy of chunk i+1 (copy stream) overlaps the
        # centring/transposition of chunk i (current stream); truly asynchronous when the
        # caller's array is pinned, still correct when it is pageable
        per = max(1, min(n, cls.CHUNK_BYTES // (12 * A)))
        dev = out.xyz.device
        src = torch.from_numpy(xyz)
        cur = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream(device=dev)
        stages = [torch.empty((per, A, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        for st in stages:
            st.record_stream(copy_stream)
        free_ev = [None, None]
        copy_stream.wait_stream(cur)
        for ci, lo in enumerate(range(0, n, per)):
            hi = min(n, lo + per)
            sidx = ci & 1
            with torch.cuda.stream(copy_stream):
                if free_ev[sidx] is not None:
                    copy_stream.wait_event(free_ev[sidx])
                stages[sidx][:hi - lo].copy_(src[lo:hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            cur.wait_event(ready)
            out.ingest_aos(stages[sidx][:hi - lo], lo, precentered)
            free_ev[sidx] = torch.cuda.Event()
            free_ev[sidx].record(cur)
        return out

    def ingest_aos(self, aos_dev, first_frame, precentered=False):
        """Centre + transpose a device-resident (m, A, 3) block into frames
        [first_frame, first_frame+m)."""
        m = int(aos_dev.shape[0])
        _lib.call("eb_center_and_trace", ptr(aos_dev), m, self.n_atoms, int(bool(precentered)),
                  ptr(self.xyz[first_frame:]), ptr(self.traces[first_frame:]), stream_ptr())

    def gather(self, idx):
        """Dense DeviceTrajectory of the frames ``idx`` (LongTensor on device or array-like)."""
        if not torch.is_tensor(idx):
            idx = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=self.xyz.device)
        idx = idx.to(torch.int64).contiguous()
        m = int(idx.numel())
        out = DeviceTrajectory.empty(m, self.n_atoms, self.top)
        _lib.call("eb_gather_frames", ptr(self.xyz), ptr(self.traces), self.n_atoms, ptr(idx),
                  m, ptr(out.xyz), ptr(out.traces), stream_ptr())
        return out

    def to_host_aos(self):
        """(n, A, 3) float32 numpy array of the CENTRED coordinates."""
        n = len(self)
        aos = torch.empty((n, self.n_atoms, 3), dtype=torch.float32, device=self.xyz.device)
        _lib.call("eb_soa_to_aos", ptr(self.xyz), n, self.n_atoms, ptr(aos), stream_ptr())
        return aos.cpu().numpy()


class DeviceFeatures:
    """Row-major (n, F) feature matrix in HBM, in the caller's dtype."""

    def __init__(self, X, host=None):
        self.X = X
        self.host = host
        self.np_dtype = np.dtype(str(X.dtype).replace("torch.", ""))
        self.dt = _NP_TO_DT[self.np_dtype]

    def __len__(self):
        return int(self.X.shape[0])

    @property
    def n_features(self):
        return int(self.X.shape[1])

    @classmethod
    def from_host(cls, X):
        arr = np.asarray(X)
        if arr.ndim != 2:
            raise DataInvalid("Data array dimension must be two, got shape %s." % (arr.shape,))
        if arr.dtype not in _NP_TO_DT:
            raise DataInvalid("Unsupported feature dtype %s (libdist accepts int8..int64, "
                              "float32, float64)." % arr.dtype)
        arr = np.ascontiguousarray(arr)
        dev = cuda_device()
        t = torch.empty(arr.shape, dtype=_NP_TO_TORCH[arr.dtype], device=dev)
        if arr.size:
            _h2d(arr, t)
        return cls(t, host=X)


def is_trajectory_like(obj):
    return hasattr(obj, "xyz") or isinstance(obj, DeviceTrajectory)
