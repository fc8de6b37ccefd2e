"""ctypes bindings to the C oracle and numpy-facing metric callables.  TEST INFRASTRUCTURE.

The callables follow the reference's metric protocol ``d = f(X, y) -> ndarray[n]``
(/root/reference/enspara/cluster/util.py:289-313, used at kcenters.py:298, kmedoids.py:637,
util.py:195,200).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libenspara_oracle.so")


def build(force=False):
    """Compile the C oracle with the Makefile next to it (gcc only, a second or two)."""
    src = os.path.join(_HERE, "enspara_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", _HERE, "-B", "libenspara_oracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, cl, ci, cd, cf = (ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double,
                              ctypes.c_float)
        L.orc_qcp_msd.restype = cd
        L.orc_qcp_msd.argtypes = [vp, cd, cd, ci]
        L.orc_center_and_trace.restype = None
        L.orc_center_and_trace.argtypes = [vp, cl, ci, vp, vp]
        L.orc_rmsd_f64.restype = None
        L.orc_rmsd_f64.argtypes = [vp, vp, cl, ci, vp, cd, vp]
        L.orc_rmsd_f32.restype = None
        L.orc_rmsd_f32.argtypes = [vp, vp, cl, ci, vp, cf, vp]
        L.orc_md_rmsd.restype = ci
        L.orc_md_rmsd.argtypes = [vp, cl, ci, vp, ci, vp]
        for sfx in ("f32", "f64", "i8", "i16", "i32", "i64"):
            for name in ("euclidean", "manhattan"):
                fn = getattr(L, "orc_%s_%s" % (name, sfx))
                fn.restype = None
                fn.argtypes = [vp, cl, cl, vp, vp]
        L.orc_kcenters_update_f32.restype = cl
        L.orc_kcenters_update_f32.argtypes = [vp, cl, cl, vp, vp]
        L.orc_num_threads.restype = ci
        L.orc_synth_trajectory.restype = None
        L.orc_synth_trajectory.argtypes = [vp, cl, ci, cl, ctypes.c_uint64, vp, ci]
        L.orc_synth_features.restype = None
        L.orc_synth_features.argtypes = [vp, cl, cl, cl, ctypes.c_uint64]
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [ci]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """Set the OpenMP team to every core this process may run on, whatever OMP_NUM_THREADS
    says (torchrun exports OMP_NUM_THREADS=1, which silently made the CPU arm single-threaded
    at N > 1).  Returns the thread count now in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(int(n))
    return num_threads()


# ---------------------------------------------------------------------------------------------
# Trajectory duck type (only .xyz, len(), X[int|slice|mask|list] are used
# on data by the reference; ops.py:208-210 also calls type(data)(xyz=, topology=)).
# ---------------------------------------------------------------------------------------------
class Trajectory:
    """Stand-in for ``mdtraj.Trajectory`` exposing exactly what the clustering path touches."""

    def __init__(self, xyz, topology=None):
        xyz = np.asarray(xyz, dtype=np.float32)
        if xyz.ndim == 2:
            xyz = xyz[None]
        self.xyz = np.ascontiguousarray(xyz)
        self.top = self.topology = topology
        self._rmsd_traces = None

    @property
    def n_frames(self):
        return self.xyz.shape[0]

    @property
    def n_atoms(self):
        return self.xyz.shape[1]

    def __len__(self):
        return self.xyz.shape[0]

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return Trajectory(self.xyz[int(key)][None].copy(), self.top)
        return Trajectory(self.xyz[key].copy(), self.top)

    def center_coordinates(self):
        t64 = np.empty(len(self), np.float64)
        lib().orc_center_and_trace(_p(self.xyz), len(self), self.n_atoms, _p(t64), None)
        self._rmsd_traces = t64
        return self


def center_and_trace(xyz):
    """Return (centred float32 copy, float64 traces, float32 traces) of an (n,A,3) array."""
    xyz = np.array(xyz, dtype=np.float32, order="C", copy=True)
    n, A = xyz.shape[0], xyz.shape[1]
    t64 = np.empty(n, np.float64)
    t32 = np.empty(n, np.float32)
    lib().orc_center_and_trace(_p(xyz), n, A, _p(t64), _p(t32))
    return xyz, t64, t32


def _xyz_of(obj):
    return obj.xyz if hasattr(obj, "xyz") else np.asarray(obj)


def _make_rmsd(mode):
    def rmsd(target, reference, frame=0, atom_indices=None, parallel=True, precentered=False):
        """Restated ``mdtraj.rmsd(target, reference, frame=0)``: float32[n] in nm."""
        X = np.ascontiguousarray(_xyz_of(target), dtype=np.float32)
        R = np.asarray(_xyz_of(reference), dtype=np.float32)
        if R.ndim == 3:
            R = R[frame]
        R = np.ascontiguousarray(R)
        if X.ndim != 3 or R.shape != X.shape[1:]:
            raise ValueError("rmsd: shape mismatch %s vs %s" % (X.shape, R.shape))
        out = np.empty(X.shape[0], np.float32)
        if X.shape[0] == 0:
            return out
        rc = lib().orc_md_rmsd(_p(X), X.shape[0], X.shape[1], _p(R), mode, _p(out))
        if rc != 0:
            raise MemoryError("orc_md_rmsd failed")
        return out
    return rmsd


#: float64-accumulation "truth" -- what the CUDA path is compared with
rmsd = _make_rmsd(0)
#: mdtraj-like float32 4-atom-lane arithmetic, scalar emulation -- noise quantification
rmsd_f32 = _make_rmsd(1)
#: the same arithmetic with SSE intrinsics (bit-identical to rmsd_f32) -- timed CPU baseline
rmsd_f32_sse = _make_rmsd(2)


def rmsd_precentered(xyz, traces64, ref, ref_trace):
    """One-vs-all on already centred data (no per-call copy)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    out = np.empty(xyz.shape[0], np.float32)
    lib().orc_rmsd_f64(_p(xyz), _p(np.ascontiguousarray(traces64)), xyz.shape[0], xyz.shape[1],
                       _p(ref), float(ref_trace), _p(out))
    return out


_SFX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int8): "i8",
        np.dtype(np.int16): "i16", np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}


def _libdist(name):
    def f(X, y, out=None):
        X = np.asarray(X)
        y = np.asarray(y)
        if X.ndim != 2 or y.ndim != 1 or X.shape[1] != y.shape[0]:
            raise ValueError("shape mismatch")
        if X.dtype != y.dtype:
            raise ValueError("Buffer dtype mismatch")
        X = np.ascontiguousarray(X)
        y = np.ascontiguousarray(y)
        if out is None:
            out = np.zeros(X.shape[0], np.float64)
        getattr(lib(), "orc_%s_%s" % (name, _SFX[X.dtype]))(
            _p(X), X.shape[0], X.shape[1], _p(y), _p(out))
        return out
    f.__name__ = name
    return f


#: restated enspara.geometry.libdist.euclidean / manhattan (libdist.pyx:148-183)
euclidean = _libdist("euclidean")
manhattan = _libdist("manhattan")


def sqeuclidean(X, x):
    """The squared-euclid callable of the reference's own test (test_cluster.py:509-510)."""
    return np.square(X - x).sum(axis=1)


def synth_trajectory(n, n_atoms, seed=0, first_frame=0, n_base=64):
    """enspara_b200.synth.trajectory restated in C + OpenMP (bit-identical, seconds instead of
    minutes): lets the CPU arms make their host sample without the product's CUDA library."""
    from enspara_b200 import synth          # numpy only: the 64 base conformers are tiny
    base = np.ascontiguousarray(synth.base_conformers(n_atoms, n_base, seed))
    out = np.empty((n, n_atoms, 3), np.float32)
    lib().orc_synth_trajectory(_p(out), n, n_atoms, first_frame, ctypes.c_uint64(seed), _p(base),
                               n_base)
    return out


def synth_features(n, n_features, seed=0, first_row=0):
    out = np.empty((n, n_features), np.float32)
    lib().orc_synth_features(_p(out), n, n_features, first_row, ctypes.c_uint64(seed))
    return out
