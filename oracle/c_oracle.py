"""ctypes wrapper of oracle/memo_oracle.c (TEST INFRASTRUCTURE; see its header)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmemo_oracle.so")


class Seg(C.Structure):
    _fields_ = [("row_begin", C.c_int64), ("n_rows", C.c_int64), ("pos0", C.c_int32),
                ("rec_len", C.c_int32), ("rec_id", C.c_int32), ("flags", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.mo_index_build.restype = C.c_int64
        _lib.mo_index_build.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(Seg),
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int64]
        _lib.mo_synth_dap.restype = None
        _lib.mo_synth_dap.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                                      C.c_uint64, C.c_int32]
        _lib.mo_query.restype = C.c_int
        _lib.mo_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    return _lib


def make_segs(records, n_rows, pos_first=0, row0=0, primed_first=True, chr_end_last=True):
    segs, acc, lo, hi = [], 0, pos_first, pos_first + n_rows
    for rid, (_, n) in enumerate(records):
        a, b = max(lo, acc), min(hi, acc + n)
        if a < b:
            segs.append(Seg(row0 + a - lo, b - a, a - acc, n, rid, 3))
        acc += n
    if acc < hi:
        raise Exception("Position beyond all intervals; ensure your fai file is from fasta of initial query.")
    if segs and not primed_first:
        segs[0].flags &= ~1
    if segs and not chr_end_last:
        segs[-1].flags &= ~2
    return segs


def index_build(dap, records, order, segs=None, cap=None):
    """dap: int32 [L, C] C-contiguous. Returns (rec, start, end, col) int64."""
    lib = load()
    dap = np.ascontiguousarray(dap, dtype=np.int32)
    L, Cc = dap.shape
    if segs is None:
        segs = make_segs(records, L)
    arr = (Seg * max(1, len(segs)))(*segs)
    if cap is None:
        cap = lib.mo_index_build(dap.ctypes.data, L, Cc, Cc, arr, len(segs), int(order), None, None, None, None, 0)
    out = [np.empty(cap, dtype=np.int64) for _ in range(4)]
    n = lib.mo_index_build(dap.ctypes.data, L, Cc, Cc, arr, len(segs), int(order),
                           out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data,
                           out[3].ctypes.data, cap)
    n = min(n, cap)
    return tuple(o[:n] for o in out)


def query(f1, f2, f3, q_start, q_end, k, n_docs, membership):
    lib = load()
    f1 = np.ascontiguousarray(f1, dtype=np.int64)
    f2 = np.ascontiguousarray(f2, dtype=np.int64)
    f3 = np.ascontiguousarray(f3, dtype=np.int64)
    W = q_end - q_start
    out = np.empty((W, n_docs), dtype=np.uint8) if membership else np.empty(W, dtype=np.int64)
    rc = lib.mo_query(f1.ctypes.data, f2.ctypes.data, f3.ctypes.data, f1.size, q_start, q_end, k,
                      n_docs, int(membership), out.ctypes.data)
    if rc != 0:
        raise IndexError("index row order/genome id out of range for -n")
    return out


def synth_dap(rec_len, n_cols, seed, row0=0, rows=None, dense=False, threads=1):
    """Rows [row0, row0+rows) of the synthetic DAP (same integers as
    memo_oracle.synth_dap), `threads` row slices generated in parallel."""
    from concurrent.futures import ThreadPoolExecutor
    lib = load()
    rows = rec_len - row0 if rows is None else rows
    out = np.empty((rows, n_cols), dtype=np.int32)
    threads = max(1, min(threads, rows // 65536 or 1))
    cuts = [(rows * i) // threads for i in range(threads + 1)]

    def work(i):
        a, b = cuts[i], cuts[i + 1]
        if b > a:
            lib.mo_synth_dap(out[a:b].ctypes.data, row0 + a, b - a, n_cols, rec_len, seed, int(dense))

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return out
