"""ctypes binding of libddrl_b200.so (C ABI declared in include/ddrl_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller
gets an exception.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libddrl_b200.so")

F32, F64, U8 = 0, 1, 2
GEMM_AUTO, GEMM_TC, GEMM_FFMA = 0, 1, 2
OK, EINVAL, ECUDA, EEMPTY, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libddrl_b200 error {code}: {msg}")
        self.code = code


_lib = None

_vp, _i64, _u64, _u32, _int = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int
_pi64, _pint = C.POINTER(C.c_int64), C.POINTER(C.c_int)
_f, _pf = C.c_float, C.POINTER(C.c_float)

# name -> (restype, argtypes).  Kept in one table so tests can check it against the header.
SIGNATURES = {
    "ddrl_abi_version": (_int, []),
    "ddrl_last_error": (C.c_char_p, []),
    "ddrl_launch_count": (_i64, []),
    "ddrl_rb_create": (_int, [_int, _int, _int, _i64, C.POINTER(_vp)]),
    "ddrl_rb_destroy": (_int, [_vp]),
    "ddrl_rb_store_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "ddrl_rb_store_batch_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "ddrl_rb_store_batch_host_copy": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "ddrl_rb_store_block_bytes": (_i64, [_vp, _i64, _int]),
    "ddrl_rb_store_block_host": (_int, [_vp, _vp, _i64, _int, _vp]),
    "ddrl_rb_sample": (_int, [_vp, _i64, _i64, _vp, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_rb_sample_host": (_int, [_vp, _i64, _i64, _vp, _u64, _u64, _u32, _vp, _i64, _vp]),
    "ddrl_rb_sample_host_async": (_int, [_vp, _i64, _i64, _vp, _u64, _u64, _u32, _vp, _i64, _vp]),
    "ddrl_rb_sample_block_bytes": (_i64, [_vp, _i64]),
    "ddrl_rb_ipc_export": (_int, [_vp, _vp]),
    "ddrl_rb_peer_attach": (_int, [_vp, _int, _int, _vp]),
    "ddrl_rb_sample_global": (_int, [_vp, _i64, _i64, _vp, _vp, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_fb_sample_stack": (_int, [_int, _vp, _i64, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _u64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_fb_store_frames": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "ddrl_seg_store": (_int, [_int, _vp, _int, _i64, _i64, _int, _pint, _pint, C.POINTER(_vp), _i64, _vp]),
    "ddrl_seg_sample": (_int, [_int, _vp, _int, _i64, _int, _pint, _pint, C.POINTER(_vp), _i64, _vp, _u64, _u64, _u32, _vp, _vp]),
    "ddrl_rb_counts": (_int, [_vp, _pi64, _pi64, _pi64, _pi64, _pi64]),
    "ddrl_rb_layout": (_int, [_vp, _pint, _pint, _pint, C.POINTER(_vp)]),
    "ddrl_rb_export": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_rb_import": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "ddrl_sac_create": (_int, [_int, _int, _int, _int, _int, _int, _f, _f, _f, _f, _f, _int, C.POINTER(_vp)]),
    "ddrl_sac_destroy": (_int, [_vp]),
    "ddrl_sac_param_count": (_i64, [_vp]),
    "ddrl_sac_set_weights": (_int, [_vp, _vp, _int, _vp]),
    "ddrl_sac_get_weights": (_int, [_vp, _vp, _int, _vp]),
    "ddrl_sac_step": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_sac_compute_grads": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _u64, _f, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_sac_grad_buffer": (_int, [_vp, C.POINTER(_vp), _pi64, C.POINTER(_vp)]),
    "ddrl_sac_apply_grads": (_int, [_vp, _int, _vp]),
    "ddrl_sac_step_dp": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _u64, _f, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_sac_step_from_buffer": (_int, [_vp, _vp, _int, _u64, _u64, _u32, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_sac_step_host": (_int, [_vp, _vp, _int, _u64, _vp, _vp, _vp, _vp, _vp]),
    "ddrl_rb_note_samples": (_int, [_vp, _i64]),
    "ddrl_rb_read_begin": (_int, [_vp, _vp, _pi64]),
    "ddrl_rb_read_end": (_int, [_vp, _vp, _i64]),
    "ddrl_sac_comm_export": (_int, [_vp, _vp]),
    "ddrl_sac_comm_attach": (_int, [_vp, _int, _int, _vp]),
    "ddrl_sac_comm_bytes": (_i64, [_vp]),
    "ddrl_sac_comm_attach_ptrs": (_int, [_vp, _int, _int, C.POINTER(_vp), _vp]),
    "ddrl_sac_comm_error": (_int, [_vp, _pint]),
    "ddrl_sac_dp_trace": (_int, [_vp, C.POINTER(C.c_uint64)]),
    "ddrl_ql_create": (_int, [_int, _int, _int, _int, _int, _int, _int, _f, _f, _f, _f, C.POINTER(_vp)]),
    "ddrl_ql_destroy": (_int, [_vp]),
    "ddrl_ql_param_count": (_i64, [_vp]),
    "ddrl_ql_set_weights": (_int, [_vp, _vp, _int, _vp]),
    "ddrl_ql_get_weights": (_int, [_vp, _vp, _int, _vp]),
    "ddrl_ql_step": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "ddrl_ql_forward": (_int, [_vp, _vp, _int, _int, _vp, _vp]),
    "ddrl_sac_act": (_int, [_vp, _vp, _int, _int, _vp, _u64, _u64, _vp, _vp]),
    "ddrl_sac_debug_stage": (_int, [_vp, _int, _int, _int, _vp]),
    "ddrl_debug_tc_gemm": (_int, [_int, _vp, _int, _int, _int, _vp, _int, _int, _int, _vp, _int, _int, _int, _int, _int, _vp]),
    "ddrl_sac_trace_stage": (_int, [_vp, _int, _int, _vp, _int, _pint, _vp]),
    "ddrl_sac_state": (_int, [_vp, _pint, _pint, _pint, _pf, _vp]),
}


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing — build it with `python distributed-drl_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(rc, lib().ddrl_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().ddrl_launch_count())
