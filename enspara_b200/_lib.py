"""ctypes binding of libenspara_b200.so (the C ABI in include/enspara_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

from .exception import DataInvalid

_HERE = os.path.dirname(os.path.abspath(__file__))
#: ENSPARA_B200_LIB points at another build of the same library (developer A/B runs only)
LIB_PATH = os.environ.get("ENSPARA_B200_LIB") or os.path.join(_HERE, "libenspara_b200.so")

EB_OK, EB_ERR_INVALID, EB_ERR_CUDA, EB_ERR_LIMIT = 0, 1, 2, 3
DT_F32, DT_F64, DT_I8, DT_I16, DT_I32, DT_I64 = range(6)
METRIC_EUCLIDEAN, METRIC_MANHATTAN, METRIC_SQEUCLIDEAN = range(3)


class KcState(ctypes.Structure):
    """Mirror of ``eb_kc_state`` (64 bytes)."""
    _fields_ = [("n_centers", ctypes.c_int32), ("done", ctypes.c_int32),
                ("blocks_done", ctypes.c_uint32), ("n_noop", ctypes.c_int32),
                ("maxdist", ctypes.c_double), ("local_maxdist", ctypes.c_double),
                ("last_center", ctypes.c_int64), ("error", ctypes.c_int64),
                ("wait_ns", ctypes.c_int64), ("reserved", ctypes.c_int64)]


assert ctypes.sizeof(KcState) == 64


class PamCtx(ctypes.Structure):
    """Mirror of ``eb_pam_ctx`` (include/enspara_b200.h): the buffers of one PAM engine."""
    _fields_ = [("xyz", ctypes.c_void_p), ("traces", ctypes.c_void_p), ("n", ctypes.c_int64),
                ("frame_offset", ctypes.c_int64), ("n_atoms", ctypes.c_int32),
                ("k", ctypes.c_int32), ("medoid_xyz", ctypes.c_void_p),
                ("medoid_traces", ctypes.c_void_p), ("prop_xyz", ctypes.c_void_p),
                ("prop_traces", ctypes.c_void_p), ("prop_idx", ctypes.c_void_p),
                ("saved_xyz", ctypes.c_void_p), ("saved_traces", ctypes.c_void_p),
                ("dist", ctypes.c_void_p), ("assign", ctypes.c_void_p),
                ("new_dist", ctypes.c_void_p), ("new_assign", ctypes.c_void_p),
                ("new_ctr_dist", ctypes.c_void_p), ("cc", ctypes.c_void_p),
                ("need_idx", ctypes.c_void_p), ("need_n", ctypes.c_void_p),
                ("need_assign", ctypes.c_void_p), ("ambig_idx", ctypes.c_void_p),
                ("scal_i", ctypes.c_void_p), ("scal_d", ctypes.c_void_p),
                ("scratch", ctypes.c_void_p), ("tc_cand", ctypes.c_void_p),
                ("tc_scratch", ctypes.c_void_p), ("tc_ovf", ctypes.c_void_p),
                ("kappa", ctypes.c_double), ("use_tc", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("pin_d", ctypes.c_void_p),
                ("pin_i", ctypes.c_void_p), ("pin_o", ctypes.c_void_p),
                ("med_list", ctypes.c_void_p), ("med_list_n", ctypes.c_void_p),
                ("med_list_cap", ctypes.c_int32), ("use_list", ctypes.c_int32)]


PAM_SELECT, PAM_TRIAL, PAM_READBACK = 1, 2, 4

_vp, _i64, _i32, _int, _dbl, _sz, _u64 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int, ctypes.c_double, ctypes.c_size_t,
                                           ctypes.c_uint64)

#: name -> (restype, argtypes); must list every symbol declared in include/enspara_b200.h
SIGNATURES = {
    "eb_version": (_int, []),
    "eb_last_error": (ctypes.c_char_p, []),
    "eb_sm_count": (_int, []),
    "eb_rmsd_apad": (_int, [_int]),
    "eb_rmsd_record_bytes": (_sz, [_int]),
    "eb_feat_record_bytes": (_sz, [_i64, _int]),
    "eb_kc_partials_bytes": (_sz, []),
    "eb_center_and_trace": (_int, [_vp, _i64, _int, _int, _vp, _vp, _vp]),
    "eb_soa_to_aos": (_int, [_vp, _i64, _int, _vp, _vp]),
    "eb_gather_frames": (_int, [_vp, _vp, _int, _vp, _i64, _vp, _vp, _vp]),
    "eb_kcenters_step_rmsd": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _int, _vp, _vp, _i32,
                                     _dbl, _vp, _vp, _vp, _vp, _int, _int, _vp]),
    "eb_kcenters_step_rmsd_uses_tma": (_int, [_i64, _int]),
    "eb_exch_bytes": (_sz, [_int, _int]),
    "eb_kcenters_seed_rmsd_p2p": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _int, _int, _vp, _i32,
                                         _vp, _vp, _vp, _vp]),
    "eb_kcenters_step_rmsd_p2p": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _int, _int, _vp, _vp,
                                         _i32, _dbl, _vp, _vp, _vp, _vp, _int, _int, _vp]),
    "eb_kcenters_step_rmsd_tri": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _int, _vp, _vp, _i32,
                                         _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                         _int, _vp]),
    "eb_kcenters_seed_rmsd": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _i32, _vp, _vp, _vp,
                                     _vp]),
    "eb_rmsd_one_to_all": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _int, _vp]),
    "eb_kcenters_step_feat": (_int, [_vp, _i64, _i64, _int, _int, _i64, _vp, _int, _vp, _vp,
                                     _i32, _dbl, _vp, _vp, _vp, _vp, _int, _vp]),
    "eb_kcenters_seed_feat": (_int, [_vp, _i64, _i64, _int, _i64, _vp, _i32, _vp, _vp, _vp,
                                     _vp]),
    "eb_feat_one_to_all": (_int, [_vp, _i64, _i64, _int, _int, _vp, _vp, _vp]),
    "eb_rmsd_one_to_all_pruned": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _vp, _i32,
                                         _vp, _vp]),
    "eb_rmsd_assign": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _int,
                              _int, _vp]),
    "eb_rmsd_assign_dev": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _i32, _vp, _i64, _vp, _vp,
                                  _int, _int, _vp, _vp]),
    "eb_rmsd_assign_tc_dev": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _i32, _dbl, _vp, _int, _vp,
                                     _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "eb_tc_scratch_bytes": (_sz, [_i64, _int, _i32]),
    "eb_rmsd_assign_tc": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _i32, _dbl, _vp, _int, _vp, _vp,
                                 _vp, _vp, _vp, _int, _vp]),
    "eb_rmsd_score_lists": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp]),
    "eb_feat_assign": (_int, [_vp, _i64, _i64, _int, _int, _vp, _i32, _vp, _i64, _vp, _vp,
                              _int, _int, _vp]),
    "eb_pam_classify": (_int, [_vp, _vp, _vp, _i64, _int, _i32, _vp, _vp, _vp, _vp, _vp]),
    "eb_pam_need_list": (_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "eb_pam_scratch_bytes": (_sz, [_i64]),
    "eb_sum_squares": (_int, [_vp, _i64, _int, _vp, _vp, _vp]),
    "eb_count_members": (_int, [_vp, _i64, _i32, _vp, _vp]),
    "eb_select_member": (_int, [_vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "eb_pam_propose_rmsd": (_int, [_vp, _i32, _i64, _i64, _int, _vp]),
    "eb_pam_restore_medoid": (_int, [_vp, _i32, _vp]),
    "eb_struct_bytes": (_sz, [_int]),
    "eb_xtc_scan": (_int, [ctypes.c_char_p, _vp, _vp]),
    "eb_xtc_read": (_int, [ctypes.c_char_p, _i64, _i64, _i64, _vp, _i32, _vp, _vp]),
    "eb_synth_trajectory_aos": (_int, [_vp, _i64, _int, _i64, _u64, _vp, _int, _vp]),
    "eb_synth_features": (_int, [_vp, _i64, _i64, _i64, _u64, _vp]),
}

#: declared under ``#ifdef EB_PLANNED`` in the header; not yet exported by the library
PLANNED = set()

_lib = None


def load():
    """Load the shared library (does not touch the GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "enspara_b200: %s is missing. Build it with `python -m enspara_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if name in PLANNED:
                continue
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    """Map a C status to the reference's error convention."""
    if status == EB_OK:
        return
    msg = load().eb_last_error().decode("utf-8", "replace")
    if status == EB_ERR_INVALID:
        raise DataInvalid(msg)
    raise RuntimeError("enspara_b200 native error %d: %s" % (status, msg))


def call(name, *args):
    check(getattr(load(), name)(*args))
