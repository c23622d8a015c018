"""Device-level Python API over libmemo_b200.so.

PyTorch is used only for device buffers and streams; every computation is a
hand-written sm_100a kernel reached through the C ABI (include/memo_b200.h).
There is no CPU fallback: without a CUDA device and the built library these
functions raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (IndexOpts, MemoError, Segment, MEMO_SEG_CHR_END, MEMO_SEG_PRIMED,
                   RES_IRREGULAR, RES_N_OUT, RES_REPLAYS, RES_SLOTS)

INT32_MAX = 2147483647


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise MemoError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise MemoError(f"{name} must be contiguous")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


# --------------------------------------------------------------------------
# record runs (host logic mirroring dap_to_bed.py:20-28, :76-83, :121-134)
# --------------------------------------------------------------------------
def parse_fai(path) -> List[Tuple[str, int]]:
    """[(header, length)] -- first two whitespace columns of a .fai
    (reference: src/dap_to_bed.py:20-28)."""
    out = []
    with open(path) as fh:
        for line in fh:
            header, length, *_ = line.strip().split()
            out.append((header, int(length)))
    return out


def segments_for_rows(records: Sequence[Tuple[str, int]], pos_first: int, n_rows: int,
                      buffer_row0: int = 0, primed_first: bool = True,
                      chr_end_last: bool = True) -> List[Segment]:
    """Record runs for DAP rows at global positions [pos_first, pos_first+n_rows)
    (the `nl -v0` numbering of index.sh:83) stored from row `buffer_row0` of a
    device buffer.

    primed_first  the first run starts fresh (False: it continues the record of
                  the previous shard/chunk and row buffer_row0-1 is its halo)
    chr_end_last  the last run is followed by its chr-end rows (False when a
                  later shard/chunk continues it)
    Raises like the reference when a position lies beyond every record
    (src/dap_to_bed.py:82-83).
    """
    segs: List[Segment] = []
    if n_rows <= 0:
        return segs
    lo, hi = pos_first, pos_first + n_rows
    acc = 0
    covered = lo
    for rid, (_, length) in enumerate(records):
        start, end = acc, acc + length
        acc = end
        a, b = max(lo, start), min(hi, end)
        if a >= b:
            continue
        if length > INT32_MAX:
            raise MemoError(f"record {rid} longer than int32 positions")
        segs.append(Segment(row_begin=buffer_row0 + (a - lo), n_rows=b - a, pos0=a - start,
                            rec_len=length, rec_id=rid,
                            flags=MEMO_SEG_PRIMED | MEMO_SEG_CHR_END))
        covered = b
    if covered < hi or lo < 0:
        raise Exception("Position beyond all intervals; ensure your fai file is from fasta "
                        "of initial query.")
    if not primed_first:
        segs[0].flags &= ~MEMO_SEG_PRIMED
    if not chr_end_last:
        segs[-1].flags &= ~MEMO_SEG_CHR_END
    return segs


# --------------------------------------------------------------------------
# index build
# --------------------------------------------------------------------------
@dataclass
class IndexRows:
    """Index rows on the device, in the reference's print order."""
    start: torch.Tensor        # int32 [n]   BED f1
    end: torch.Tensor          # int64-safe uint32 stored as int32 bits [n]   BED f2
    order: torch.Tensor        # int32 [n]   BED f3
    seg_out_end: torch.Tensor  # int64 [n_seg] cumulative rows per run (host)
    seg_rec_id: List[int]
    n: int
    irregular: bool            # input was not valid matching statistics
    replays: int
    general: bool              # produced by the three-pass general build

    def to_host(self):
        """(rec_idx, start, end, order) int64 numpy arrays."""
        n = self.n
        start = self.start[:n].cpu().numpy().astype(np.int64)
        end = self.end[:n].cpu().numpy().view(np.uint32).astype(np.int64)
        order = self.order[:n].cpu().numpy().astype(np.int64)
        counts = np.diff(np.concatenate([[0], self.seg_out_end.numpy()]))
        rec = np.repeat(np.asarray(self.seg_rec_id, dtype=np.int64), counts)
        return rec, start, end, order


def _opts(order: bool, rows_per_tile=0, emit_buf_records=0, warps_per_cta=0, ctas_per_sm=0,
          stages=0, kernel_variant=0) -> IndexOpts:
    return IndexOpts(order_mode=1 if order else 0, rows_per_tile=rows_per_tile,
                     emit_buf_records=emit_buf_records, warps_per_cta=warps_per_cta,
                     ctas_per_sm=ctas_per_sm, stages=stages, kernel_variant=kernel_variant)


class IndexBuilder:
    """Reusable launcher for one DAP buffer shape (keeps workspace / outputs)."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise MemoError("no CUDA device: memo_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self._ws = None
        self._result = torch.zeros(RES_SLOTS, dtype=torch.int64, device=self.device)

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        return self._ws

    def launch(self, dap: torch.Tensor, n_cols: int, segs: Sequence[Segment], order: bool,
               out: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], seg_out_end: torch.Tensor,
               general: bool = False, carry_in: Optional[torch.Tensor] = None,
               carry_out: Optional[torch.Tensor] = None, result: Optional[torch.Tensor] = None,
               **tuning) -> None:
        """Enqueue one build on the current stream (no synchronisation).  `result`
        (device int64 [RES_SLOTS]) replaces the builder's own result slots."""
        _require_cuda(dap, "dap")
        if dap.dtype != torch.int32 or dap.dim() != 2:
            raise MemoError("dap must be int32 [rows, ld]")
        rows, ld = dap.shape
        seg_arr = (Segment * max(len(segs), 1))(*segs)
        opts = _opts(order, **tuning)
        cap = 0 if out is None else out[0].numel()
        need = self.lib.memo_index_workspace_bytes(rows, n_cols, ld, cap, seg_arr, len(segs),
                                                   C.byref(opts))
        if need == 0:
            raise MemoError("memo_index_workspace_bytes: " + self.lib.memo_last_error().decode())
        ws = self._workspace(need)
        o0, o1, o2 = (None, None, None) if out is None else out
        stream = _stream_ptr(self.device)
        res_ptr = self._result.data_ptr() if result is None else result.data_ptr()
        if general:
            rc = self.lib.memo_index_build_general(
                dap.data_ptr(), rows, n_cols, ld, seg_arr, len(segs), C.byref(opts),
                _ptr(carry_in), _ptr(carry_out), _ptr(o0), _ptr(o1), _ptr(o2), cap,
                _ptr(seg_out_end), res_ptr, ws.data_ptr(), ws.numel(), stream)
            _lib.check(rc, "memo_index_build_general")
        else:
            rc = self.lib.memo_index_build(
                dap.data_ptr(), rows, n_cols, ld, seg_arr, len(segs), C.byref(opts),
                _ptr(o0), _ptr(o1), _ptr(o2), cap, _ptr(seg_out_end),
                res_ptr, ws.data_ptr(), ws.numel(), stream)
            _lib.check(rc, "memo_index_build")

    def result(self) -> Tuple[int, bool, int]:
        r = self._result.cpu().tolist()     # synchronises the stream
        return int(r[RES_N_OUT]), bool(r[RES_IRREGULAR]), int(r[RES_REPLAYS])

    def build(self, dap: torch.Tensor, n_cols: int, segs: Sequence[Segment], order: bool,
              out_cap: Optional[int] = None, force_general: bool = False, **tuning) -> IndexRows:
        """Build the index rows for `dap`; exact for every non-negative integer
        input (falls back to the general build when the device flags the input
        as irregular; grows the output buffers when they were too small)."""
        rows = dap.shape[0]
        if out_cap is None:
            out_cap = max(1024, int(rows * n_cols * 0.02) + n_cols * (len(segs) + 1))
        seg_out_end = torch.zeros(max(len(segs), 1), dtype=torch.int64, device=self.device)
        general = force_general
        while True:
            out = tuple(torch.empty(out_cap, dtype=torch.int32, device=self.device) for _ in range(3))
            self.launch(dap, n_cols, segs, order, out, seg_out_end, general=general, **tuning)
            n, irregular, replays = self.result()
            if irregular and not general:
                general = True
                continue
            if n > out_cap:
                out_cap = n
                continue
            break
        return IndexRows(start=out[0], end=out[1], order=out[2],
                         seg_out_end=seg_out_end[:len(segs)].cpu(),
                         seg_rec_id=[s.rec_id for s in segs], n=n, irregular=irregular,
                         replays=replays, general=general)


def index_build(dap: torch.Tensor, records: Sequence[Tuple[str, int]], order: bool,
                n_cols: Optional[int] = None, **kw) -> IndexRows:
    """DAP (device int32 [L, C], rows = global positions 0..L-1) -> index rows."""
    n_cols = dap.shape[1] if n_cols is None else n_cols
    segs = segments_for_rows(records, 0, dap.shape[0])
    return IndexBuilder(dap.device).build(dap, n_cols, segs, order, **kw)


# --------------------------------------------------------------------------
# query
# --------------------------------------------------------------------------
def _query_common(f1, f2, f3):
    for t, n in ((f1, "f1"), (f2, "f2"), (f3, "f3")):
        _require_cuda(t, n)
        if t.dtype != torch.int32 or t.dim() != 1:
            raise MemoError(f"{n} must be int32 [n_rows]")
    if not (f1.numel() == f2.numel() == f3.numel()):
        raise MemoError("f1/f2/f3 length mismatch")


def query_conservation(f1: torch.Tensor, f2: torch.Tensor, f3: torch.Tensor, q_start: int,
                       q_end: int, k: int, n_docs: int, out: Optional[torch.Tensor] = None,
                       check: bool = True, status: Optional[torch.Tensor] = None,
                       workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conservation vector (uint8, or int16-typed uint16 when n_docs > 255) for the
    window [q_start, q_end).  f1/f2/f3: one record's index rows, f1 ascending
    (f2 holds uint32 bits)."""
    lib = _lib.load()
    _query_common(f1, f2, f3)
    dev = f1.device
    W = max(0, q_end - q_start)
    u16 = n_docs > 255
    if out is None:
        out = torch.empty(W, dtype=torch.int16 if u16 else torch.uint8, device=dev)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    need = lib.memo_query_workspace_bytes(W)
    ws = workspace if workspace is not None and workspace.numel() >= need else \
        torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    rc = lib.memo_query_conservation(_ptr(f1), _ptr(f2), _ptr(f3), f1.numel(), q_start, q_end, k,
                                     n_docs, _ptr(out), 1 if u16 else 0, status.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "memo_query_conservation")
    if check and int(status.item()) != 0:
        raise IndexError("index row order out of range for -n (the reference writes out of bounds)")
    return out


def query_membership(f1: torch.Tensor, f2: torch.Tensor, f3: torch.Tensor, q_start: int,
                     q_end: int, k: int, n_docs: int, out: Optional[torch.Tensor] = None,
                     check: bool = True, status: Optional[torch.Tensor] = None,
                     workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Membership bitmaps: int32-typed uint32 [W, ceil(n_docs/32)]."""
    lib = _lib.load()
    _query_common(f1, f2, f3)
    dev = f1.device
    W = max(0, q_end - q_start)
    nw = (n_docs + 31) // 32
    if out is None:
        out = torch.empty((W, nw), dtype=torch.int32, device=dev)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    need = lib.memo_query_workspace_bytes(W)
    ws = workspace if workspace is not None and workspace.numel() >= need else \
        torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    rc = lib.memo_query_membership(_ptr(f1), _ptr(f2), _ptr(f3), f1.numel(), q_start, q_end, k,
                                   n_docs, _ptr(out), status.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream_ptr(dev))
    _lib.check(rc, "memo_query_membership")
    if check and int(status.item()) != 0:
        raise IndexError("index row genome id out of range for -n (the reference writes out of bounds)")
    return out


def query_sweep(f1: torch.Tensor, f2: torch.Tensor, f3: torch.Tensor, q_start: int, q_end: int,
                ks: Sequence[int], n_docs: int, membership: bool = False,
                workspace: Optional[torch.Tensor] = None, check: bool = True) -> torch.Tensor:
    """The window [q_start, q_end) answered for every k of `ks` (BASELINE configs[4]: 15 .. 101) in
    one launch per 16 values: conservation uint8 [len(ks), W16] (W16 = W rounded up to 16; row i =
    the vector for ks[i] in its first W bytes) or membership int32-typed uint32 [len(ks), W, NW]."""
    lib = _lib.load()
    _query_common(f1, f2, f3)
    dev = f1.device
    W = max(0, q_end - q_start)
    ks = [int(k) for k in ks]
    if membership:
        nw = (n_docs + 31) // 32
        if (W * nw) % 4:
            raise MemoError("membership sweep: W * ceil(n_docs / 32) must be a multiple of 4")
        out = torch.empty((len(ks), W, nw), dtype=torch.int32, device=dev)
    else:
        out = torch.empty((len(ks), (W + 15) // 16 * 16), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    need = lib.memo_query_workspace_bytes(W)
    ws = workspace if workspace is not None and workspace.numel() >= need else \
        torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    for a in range(0, len(ks), 16):
        part = ks[a:a + 16]
        arr = (C.c_int32 * len(part))(*part)
        rc = lib.memo_query_sweep(1 if membership else 0, _ptr(f1), _ptr(f2), _ptr(f3), f1.numel(), q_start, q_end,
                                  arr, len(part), n_docs, out[a].data_ptr() if W else 0, status.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(rc, "memo_query_sweep")
        bad |= status
    if check and int(bad.item()) != 0:
        raise IndexError("index row order / genome id out of range for -n (the reference writes out of bounds)")
    return out


def unpack_membership(bits: np.ndarray, n_docs: int) -> np.ndarray:
    """uint32 [W, NW] -> uint8 [W, n_docs] (host helper for tests / writers)."""
    b = np.ascontiguousarray(bits).view(np.uint32)
    W = b.shape[0]
    cols = np.arange(n_docs)
    return ((b[:, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(np.uint8).reshape(W, n_docs)


# --------------------------------------------------------------------------
# text formatters
# --------------------------------------------------------------------------
def format_conservation(vals: torch.Tensor) -> bytes:
    """Device conservation vector -> the bytes memo_query.py:71 writes."""
    lib = _lib.load()
    _require_cuda(vals, "vals")
    n = vals.numel()
    u16 = vals.dtype == torch.int16
    dev = vals.device
    text = torch.empty(max(1, n * (6 if u16 else 4)), dtype=torch.uint8, device=dev)
    out_len = torch.zeros(1, dtype=torch.int64, device=dev)
    need = lib.memo_format_workspace_bytes(n)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    rc = lib.memo_format_conservation(_ptr(vals), 1 if u16 else 0, n, text.data_ptr(),
                                      out_len.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "memo_format_conservation")
    m = int(out_len.item())
    return text[:m].cpu().numpy().tobytes()


class BedFormatter:
    """Index rows on the device -> the BED text dap_to_bed.py prints (src/dap_to_bed.py:100-109),
    formatted on the device and copied back as bytes.  Keeps its device / pinned buffers."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise MemoError("no CUDA device: memo_b200 has no CPU fallback")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.text = self.ws = self.pin = None
        self.out_len = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.pin_len = torch.zeros(1, dtype=torch.int64, pin_memory=True)

    def format(self, rows: torch.Tensor, name: str):
        """rows: device int32 [3, n] (start, end as uint32 bits, order) of ONE record ->
        uint8 numpy view of the text (valid until the next call)."""
        if not rows.is_cuda:
            raise MemoError("rows must be a CUDA tensor (no CPU fallback)")
        if rows.dtype != torch.int32 or rows.dim() != 2 or rows.shape[0] != 3 or \
                (rows.shape[1] > 1 and rows.stride(1) != 1):
            raise MemoError("rows must be int32 [3, n] with contiguous columns")
        n = rows.shape[1]
        nm = name.encode("utf-8")
        need_text = self.lib.memo_format_bed_max_bytes(n, len(nm))
        if self.text is None or self.text.numel() < need_text:
            self.text = torch.empty(max(need_text, 1 << 16), dtype=torch.uint8, device=self.dev)
        need_ws = self.lib.memo_format_bed_workspace_bytes(n)
        if self.ws is None or self.ws.numel() < need_ws:
            self.ws = torch.empty(max(need_ws, 1 << 12), dtype=torch.uint8, device=self.dev)
        rc = self.lib.memo_format_bed(rows[0].data_ptr(), rows[1].data_ptr(), rows[2].data_ptr(), n, nm, len(nm),
                                      self.text.data_ptr(), self.out_len.data_ptr(), self.ws.data_ptr(),
                                      self.ws.numel(), _stream_ptr(self.dev))
        _lib.check(rc, "memo_format_bed")
        self.pin_len.copy_(self.out_len, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        m = int(self.pin_len[0])
        if self.pin is None or self.pin.numel() < m:
            self.pin = torch.empty(max(m + m // 4, 1 << 16), dtype=torch.uint8, pin_memory=True)
        self.pin[:m].copy_(self.text[:m], non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.pin.numpy()[:m]


def format_membership(bits: torch.Tensor, n_docs: int) -> bytes:
    """Device membership bitmaps -> the bytes np.savetxt writes (memo_query.py:68)."""
    lib = _lib.load()
    _require_cuda(bits, "bits")
    W = bits.shape[0]
    dev = bits.device
    text = torch.empty(max(2, 2 * W * n_docs), dtype=torch.uint8, device=dev)
    rc = lib.memo_format_membership(_ptr(bits), W, n_docs, text.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "memo_format_membership")
    return text[:2 * W * n_docs].cpu().numpy().tobytes()


# --------------------------------------------------------------------------
# dap.txt text on the device
# --------------------------------------------------------------------------
class DapTextParser:
    """dap.txt bytes -> int32 DAP rows on the device (memo_dap_text_parse; replaces the line
    reader and `map(int, row.split(' '))` of src/dap_to_bed.py:14-18,85-88).  Reusable: keeps
    its device buffers (text, rows, workspace) sized for `block_bytes` of text."""

    ERRORS = ((1, ValueError, "invalid literal for int() in dap.txt (a character that is no digit, single "
                              "space or newline, or an empty field)"),
              (2, MemoError, "dap.txt: a line with another number of fields than the first line"),
              (4, MemoError, "dap.txt positions are not consecutive (expected `nl -v0` numbering)"),
              (8, MemoError, "DAP lengths must be in [0, 2^31)"),
              (16, MemoError, "dap.txt: more lines in a block than its size allows"))

    def __init__(self, n_cols: int, block_bytes: int, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise MemoError("no CUDA device: memo_b200 has no CPU fallback")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.C, self.block_bytes = n_cols, int(block_bytes)
        self.text = torch.empty(self.block_bytes + 64, dtype=torch.uint8, device=self.dev)
        # a line has at least 2 bytes per field ("0 ")
        self.max_rows = self.block_bytes // (2 * (n_cols + 1)) + 2
        self.rows = torch.empty((self.max_rows, n_cols), dtype=torch.int32, device=self.dev)
        self.result = torch.zeros(4, dtype=torch.int64, device=self.dev)
        self.ws = torch.empty(max(self.lib.memo_dap_text_workspace_bytes(self.block_bytes), 1), dtype=torch.uint8,
                              device=self.dev)

    def parse(self, host_bytes: torch.Tensor, n_bytes: int, pos_first: int) -> torch.Tensor:
        """host_bytes: pinned uint8 tensor holding n_bytes of WHOLE lines.  Returns the device
        rows int32 [n_lines, C] (a view of this parser's buffer: valid until the next parse)."""
        if n_bytes > self.block_bytes:
            raise MemoError("text block larger than the parser was sized for")
        self.text[:n_bytes].copy_(host_bytes[:n_bytes], non_blocking=True)
        rc = self.lib.memo_dap_text_parse(self.text.data_ptr(), n_bytes, self.C, pos_first, self.rows.data_ptr(),
                                          self.max_rows, self.C, self.result.data_ptr(), self.ws.data_ptr(),
                                          self.ws.numel(), _stream_ptr(self.dev))
        _lib.check(rc, "memo_dap_text_parse")
        n_lines, err = self.result[:2].tolist()             # synchronises
        for bit, exc, msg in self.ERRORS:
            if err & bit:
                raise exc(msg)
        return self.rows[:n_lines]


# --------------------------------------------------------------------------
# view binning
# --------------------------------------------------------------------------
def view_bin_edges(n_positions: int, n_bins: int) -> np.ndarray:
    """Bin edges exactly as src/plot_conservation.py:52: int(linspace(0, positions, n_bins + 1))."""
    return np.array(list(map(int, np.linspace(0, n_positions, n_bins + 1))), dtype=np.int64)


def view_bins(vals: torch.Tensor, n_docs: int, n_bins: int) -> np.ndarray:
    """Per-bin counts of the conservation values 0 .. n_docs (uint64 numpy [n_bins, n_docs + 1])
    of a device conservation vector (src/plot_conservation.py:46-58, the Counter per bin)."""
    lib = _lib.load()
    _require_cuda(vals, "vals")
    if vals.dtype not in (torch.uint8, torch.int16):
        raise MemoError("vals must be a uint8 / int16-typed uint16 conservation vector")
    dev = vals.device
    n = vals.numel()
    edges = torch.from_numpy(view_bin_edges(n, n_bins)).to(dev)
    counts = torch.empty((n_bins, n_docs + 1), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.memo_view_bins(_ptr(vals), 1 if vals.dtype == torch.int16 else 0, n, n_docs, n_bins,
                            edges.data_ptr(), counts.data_ptr(), status.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "memo_view_bins")
    # (values above n_docs are not in the table but stay in the bin sizes, as in the reference)
    return counts.cpu().numpy().view(np.uint64)


# --------------------------------------------------------------------------
# synthetic DAP
# --------------------------------------------------------------------------
def synth_dap(rec_len: int, n_cols: int, seed: int, row0: int = 0, rows: Optional[int] = None,
              dense: bool = False, device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows [row0, row0+rows) of the synthetic HPRC-shaped DAP of one record
    (bit-identical to oracle.memo_oracle.synth_dap), generated on the device."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise MemoError("no CUDA device: memo_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rows = rec_len - row0 if rows is None else rows
    if out is None:
        out = torch.empty((rows, n_cols), dtype=torch.int32, device=dev)
    need = lib.memo_synth_workspace_bytes(rows, n_cols)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    rc = lib.memo_synth_dap(out.data_ptr(), row0, rows, n_cols, out.stride(0), rec_len, seed,
                            1 if dense else 0, ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "memo_synth_dap")
    return out
