"""ctypes binding of libmemo_b200.so (include/memo_b200.h).

There is no CPU fallback: if the library has not been built (or cannot be
loaded) every device entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmemo_b200.so")

MEMO_SEG_PRIMED = 1
MEMO_SEG_CHR_END = 2
RES_N_OUT, RES_IRREGULAR, RES_REPLAYS, RES_SLOTS = 0, 1, 2, 4


class Segment(C.Structure):
    """memo_segment_t"""
    _fields_ = [("row_begin", C.c_int64), ("n_rows", C.c_int64), ("pos0", C.c_int32),
                ("rec_len", C.c_int32), ("rec_id", C.c_int32), ("flags", C.c_int32)]


class IndexOpts(C.Structure):
    """memo_index_opts_t"""
    _fields_ = [("order_mode", C.c_int32), ("rows_per_tile", C.c_int32),
                ("emit_buf_records", C.c_int32), ("warps_per_cta", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("stages", C.c_int32),
                ("kernel_variant", C.c_int32), ("reserved", C.c_int32)]


class LengthsFile(C.Structure):
    """memo_lengths_file_t"""
    _fields_ = [("fd", C.c_int32), ("state", C.c_int32), ("eof", C.c_int32), ("ended", C.c_int32),
                ("buf", C.c_void_p), ("cap", C.c_int64), ("lo", C.c_int64), ("hi", C.c_int64),
                ("count", C.c_int64), ("error", C.c_int64)]


# symbol -> (restype, argtypes); must list every function include/memo_b200.h declares
_vp, _i32, _i64, _sz, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_uint64
SIGNATURES = {
    "memo_abi_version": (C.c_int, []),
    "memo_last_error": (C.c_char_p, []),
    "memo_device_sm_count": (C.c_int, []),
    "memo_launch_count": (_i64, [_i32]),
    "memo_index_workspace_bytes": (_sz, [_i64, _i32, _i32, _i64, C.POINTER(Segment), _i32,
                                         C.POINTER(IndexOpts)]),
    "memo_index_build": (C.c_int, [_vp, _i64, _i32, _i32, C.POINTER(Segment), _i32,
                                   C.POINTER(IndexOpts), _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "memo_index_build_general": (C.c_int, [_vp, _i64, _i32, _i32, C.POINTER(Segment), _i32,
                                           C.POINTER(IndexOpts), _vp, _vp, _vp, _vp, _vp, _i64,
                                           _vp, _vp, _vp, _sz, _vp]),
    "memo_profile_enable": (C.c_int, [_i32]),
    "memo_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "memo_query_workspace_bytes": (_sz, [_i64]),
    "memo_query_conservation": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _i32,
                                          _vp, _vp, _sz, _vp]),
    "memo_query_membership": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp,
                                        _vp, _sz, _vp]),
    "memo_query_sweep": (C.c_int, [_i32, _vp, _vp, _vp, _i64, _i64, _i64, C.POINTER(C.c_int32), _i32, _i32, _vp,
                                   _vp, _vp, _sz, _vp]),
    "memo_synth_workspace_bytes": (_sz, [_i64, _i32]),
    "memo_synth_dap": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i64, _u64, _i32, _vp, _sz, _vp]),
    "memo_format_bed_workspace_bytes": (_sz, [_i64]),
    "memo_format_bed_max_bytes": (_sz, [_i64, _i32]),
    "memo_format_bed": (C.c_int, [_vp, _vp, _vp, _i64, C.c_char_p, _i32, _vp, _vp, _vp, _sz, _vp]),
    "memo_format_workspace_bytes": (_sz, [_i64]),
    "memo_format_conservation": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _sz, _vp]),
    "memo_format_membership": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "memo_dap_text_workspace_bytes": (_sz, [_i64]),
    "memo_dap_text_parse": (C.c_int, [_vp, _i64, _i32, _i64, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "memo_lengths_text_parse": (C.c_int, [_vp, _i64, _i32, C.POINTER(C.c_int32), _vp, _i64, _i64,
                                          C.POINTER(C.c_int64)]),
    "memo_lengths_block_parse": (C.c_int, [C.POINTER(LengthsFile), _i32, _vp, _i64, _i64, _i64]),
    "memo_view_bins": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
}

_lib = None


class MemoError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MemoError(
            f"{LIB_PATH} is missing: build it with `python -m memo_b200._build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().memo_last_error().decode("utf-8", "replace")
        raise MemoError(f"{what} failed (code {rc}): {msg}")
