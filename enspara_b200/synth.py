"""Synthetic inputs of the shapes BASELINE.json names, from a counter-based
generator keyed on (seed, global frame index): any shard of any size is reproducible anywhere,
and the CUDA generator (csrc/eb_synth.cu) produces bit-identical values because only
integer hashing and correctly rounded float32 + - * / sqrt are used (no transcendental
functions, no fused multiply-adds).

Trajectories: ``n_base`` base conformers (3-D random walks with 0.38 nm steps); frame f is
base[f mod n_base] + sigma_f * noise, sigma_f ~ U(0.02, 0.15) nm, followed by a random proper
rotation and a translation in [-1, 1)^3 nm, so centring and superposition are both exercised.
Features: U[0, 1) float32.
"""
import numpy as np

_U64 = np.uint64
_GOLD = _U64(0x9E3779B97F4A7C15)
_M1 = _U64(0xBF58476D1CE4E5B9)
_M2 = _U64(0x94D049BB133111EB)
_M3 = _U64(0xD1B54A32D192ED03)
_F32 = np.float32
_SQRT3 = _F32(1.7320508)
_TWO_M24 = _F32(2.0 ** -24)


def _mix(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = np.asarray(z, dtype=_U64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> _U64(30))) * _M1
        z = (z ^ (z >> _U64(27))) * _M2
        return z ^ (z >> _U64(31))


def _stream_key(seed, stream):
    with np.errstate(over="ignore"):
        return _mix(_U64(seed) + _GOLD * _U64(stream + 1))


def u01(seed, stream, idx, counter):
    """float32 uniform in [0,1) with 24 random bits for (stream, idx, counter)."""
    idx = np.asarray(idx, dtype=_U64)
    counter = np.asarray(counter, dtype=_U64)
    with np.errstate(over="ignore"):
        z = _mix(_stream_key(seed, stream) ^ _mix(idx * _M3 + counter))
    return (z >> _U64(40)).astype(_F32) * _TWO_M24


def gauss4(seed, stream, idx, counter4):
    """Approximately N(0,1): centred sum of four uniforms times sqrt(3); counters
    counter4 .. counter4+3 are consumed."""
    c = np.asarray(counter4, dtype=_U64)
    u0 = u01(seed, stream, idx, c)
    u1 = u01(seed, stream, idx, c + _U64(1))
    u2 = u01(seed, stream, idx, c + _U64(2))
    u3 = u01(seed, stream, idx, c + _U64(3))
    return (((u0 + u1) + (u2 + u3)) - _F32(2.0)) * _SQRT3


STREAM_BASE, STREAM_FRAME, STREAM_FEAT = 0, 1, 2


def base_conformers(n_atoms, n_base=64, seed=0, step=0.38):
    """(n_base, n_atoms, 3) float32 random walks; generated on the host also for the CUDA path
    (they are tiny) and handed to eb_synth_trajectory_aos."""
    b = np.arange(n_base, dtype=np.uint64)[:, None, None]
    a = np.arange(n_atoms, dtype=np.uint64)[None, :, None]
    c = np.arange(3, dtype=np.uint64)[None, None, :]
    g = gauss4(seed, STREAM_BASE, b, (a * _U64(3) + c) * _U64(4))
    norm = np.sqrt((g[..., 0] * g[..., 0] + g[..., 1] * g[..., 1]) + g[..., 2] * g[..., 2])
    norm = np.maximum(norm, _F32(1e-6))
    steps = g / norm[..., None] * _F32(step)
    return np.cumsum(steps.astype(_F32), axis=1, dtype=_F32)


def trajectory(n, n_atoms, seed=0, first_frame=0, n_base=64, base=None):
    """(n, n_atoms, 3) float32 frames first_frame .. first_frame+n of the synthetic trajectory."""
    if base is None:
        base = base_conformers(n_atoms, n_base, seed)
    n_base = base.shape[0]
    f = np.arange(first_frame, first_frame + n, dtype=np.uint64)
    sigma = _F32(0.02) + _F32(0.13) * u01(seed, STREAM_FRAME, f, 0)
    # rotation from a normalised quaternion (counters 1..16)
    q = np.stack([gauss4(seed, STREAM_FRAME, f, 1 + 4 * j) for j in range(4)], axis=1)
    qn = np.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + (q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3]))
    qn = np.maximum(qn, _F32(1e-6))
    w, x, y, z = (q[:, j] / qn for j in range(4))
    two = _F32(2.0)
    one = _F32(1.0)
    R = np.empty((n, 3, 3), _F32)
    R[:, 0, 0] = one - two * (y * y + z * z)
    R[:, 0, 1] = two * (x * y - z * w)
    R[:, 0, 2] = two * (x * z + y * w)
    R[:, 1, 0] = two * (x * y + z * w)
    R[:, 1, 1] = one - two * (x * x + z * z)
    R[:, 1, 2] = two * (y * z - x * w)
    R[:, 2, 0] = two * (x * z - y * w)
    R[:, 2, 1] = two * (y * z + x * w)
    R[:, 2, 2] = one - two * (x * x + y * y)
    t = np.stack([two * u01(seed, STREAM_FRAME, f, 17 + c) - one for c in range(3)], axis=1)
    # per-atom noise: counters 32 + (3a+c)*4 ..
    a = np.arange(n_atoms, dtype=np.uint64)[None, :, None]
    c = np.arange(3, dtype=np.uint64)[None, None, :]
    noise = gauss4(seed, STREAM_FRAME, f[:, None, None], _U64(32) + (a * _U64(3) + c) * _U64(4))
    p = base[(f % _U64(n_base)).astype(np.int64)] + sigma[:, None, None] * noise
    px, py, pz = p[..., 0], p[..., 1], p[..., 2]
    out = np.empty((n, n_atoms, 3), _F32)
    for i in range(3):
        out[..., i] = ((R[:, i, 0, None] * px + R[:, i, 1, None] * py)
                       + R[:, i, 2, None] * pz) + t[:, i, None]
    return out


def features(n, n_features, seed=0, first_row=0):
    """(n, n_features) float32 U[0,1)."""
    r = np.arange(first_row, first_row + n, dtype=np.uint64)[:, None]
    j = np.arange(n_features, dtype=np.uint64)[None, :]
    return u01(seed, STREAM_FEAT, r, j)


# ---------------------------------------------------------------------------------------------
# device-side generation (csrc/eb_synth.cu): same values, written straight into HBM
# ---------------------------------------------------------------------------------------------
def device_trajectory_aos(n, n_atoms, seed=0, first_frame=0, n_base=64, base=None):
    """(n, n_atoms, 3) float32 CUDA tensor, bit-identical to ``trajectory(...)``."""
    import ctypes

    import torch

    from . import _lib
    from .device import cuda_device, ptr, stream_ptr
    if base is None:
        base = base_conformers(n_atoms, n_base, seed)
    dev = cuda_device()
    base_dev = torch.from_numpy(np.ascontiguousarray(base)).to(dev)
    out = torch.empty((n, n_atoms, 3), dtype=torch.float32, device=dev)
    _lib.call("eb_synth_trajectory_aos", ptr(out), int(n), int(n_atoms), int(first_frame),
              ctypes.c_uint64(int(seed)), ptr(base_dev), int(base.shape[0]), stream_ptr())
    return out


def device_trajectory(n, n_atoms, seed=0, first_frame=0, n_base=64, chunk_frames=None):
    """A ``DeviceTrajectory`` (centred SoA + traces) of synthetic frames generated in HBM chunk
    by chunk; the raw AoS frames only ever exist for one chunk at a time."""
    from .device import DeviceTrajectory
    base = base_conformers(n_atoms, n_base, seed)
    out = DeviceTrajectory.empty(n, n_atoms)
    if chunk_frames is None:
        chunk_frames = max(1, (1 << 29) // (12 * n_atoms))
    for lo in range(0, n, chunk_frames):
        m = min(chunk_frames, n - lo)
        aos = device_trajectory_aos(m, n_atoms, seed, first_frame + lo, base=base)
        out.ingest_aos(aos, lo)
    return out


def device_features(n, n_features, seed=0, first_row=0):
    """(n, n_features) float32 ``DeviceFeatures``, bit-identical to ``features(...)``."""
    import ctypes

    import torch

    from . import _lib
    from .device import DeviceFeatures, cuda_device, ptr, stream_ptr
    X = torch.empty((n, n_features), dtype=torch.float32, device=cuda_device())
    _lib.call("eb_synth_features", ptr(X), int(n), int(n_features), int(first_row),
              ctypes.c_uint64(int(seed)), stream_ptr())
    return DeviceFeatures(X)
