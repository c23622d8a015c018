"""Host-buffer API: the calls the drop-in scripts (and bench.py's e2e leg) make.

Inputs and outputs are HOST arrays; host<->device copies happen inside.

`build_index` streams the DAP through a small ring of device buffers: chunk i+1
is copied (pinned host -> device, copy stream) while the build kernels of chunk
i run (compute stream).  Every chunk carries the row before it as a halo, so
chunks are independent position ranges of the pivot; their index rows are
concatenated in chunk order, which is the reference's print order.  Nothing is
synchronised until all chunks are enqueued.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import api
from ._lib import MemoError, RES_IRREGULAR, RES_N_OUT, RES_SLOTS

DEFAULT_CHUNK_BYTES = 256 << 20
RING = 3

_pinned_pool = {}
_owned_blocks = []          # [(pinned uint8 tensor, weakref to the numpy array handed out)]


def _pinned(tag: str, nbytes: int) -> torch.Tensor:
    """Reusable pinned staging buffer (cudaHostAlloc is slow; keep them)."""
    buf = _pinned_pool.get(tag)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, pin_memory=True)
        _pinned_pool[tag] = buf
    return buf


def _pinned_result(nbytes: int):
    """Pinned block for a RESULT that is handed to the caller without another
    copy.  Returns (uint8 tensor view, numpy uint8 array over the same bytes);
    the block is recycled once the caller has dropped every view of the array."""
    best = None
    for i, (blk, ref) in enumerate(_owned_blocks):
        if ref is not None and ref() is None and blk.numel() >= nbytes and (best is None or blk.numel() < _owned_blocks[best][0].numel()):
            best = i
    if best is None:
        blk = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, pin_memory=True)
        _owned_blocks.append((blk, None))
        best = len(_owned_blocks) - 1
        # keep the pool small: forget free blocks beyond a handful
        free = [i for i, (b, r) in enumerate(_owned_blocks) if r is not None and r() is None and i != best]
        for i in sorted(free[4:], reverse=True):
            del _owned_blocks[i]
            if i < best:
                best -= 1
    blk = _owned_blocks[best][0]
    root = blk.numpy()                      # every view handed out has this array as its base
    _owned_blocks[best] = (blk, weakref.ref(root))
    return blk[:nbytes], root[:nbytes]


def _as_host_tensor(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.is_cuda:
        raise MemoError("host API expects host buffers")
    return t.contiguous()


def _clip_segments(segs_all: List[api.Segment], a: int, b: int, shift: int) -> List[api.Segment]:
    """Record runs restricted to buffer rows [a, b), re-based so that row a sits at
    row `shift` of the chunk buffer: a run cut on the left continues its record
    (halo = the row before it), a run cut on the right does not emit its chr-end
    rows yet."""
    out = []
    for s in segs_all:
        lo, hi = max(a, s.row_begin), min(b, s.row_begin + s.n_rows)
        if lo >= hi:
            continue
        flags = s.flags
        if lo > s.row_begin:
            flags &= ~api.MEMO_SEG_PRIMED
        if hi < s.row_begin + s.n_rows:
            flags &= ~api.MEMO_SEG_CHR_END
        out.append(api.Segment(row_begin=lo - a + shift, n_rows=hi - lo, pos0=s.pos0 + (lo - s.row_begin),
                               rec_len=s.rec_len, rec_id=s.rec_id, flags=flags))
    return out


class IndexRowsHost:
    """Index rows on the host in the reference's print order.  start / end /
    order are int32 / uint32 / int32 arrays (end = BED f2 may exceed int32);
    `as_int64()` gives the (rec_idx, start, end, order) int64 tuple."""

    def __init__(self, rec_counts, start, end, order, general):
        self.rec_counts = rec_counts          # [(rec_id, n_rows)] in output order
        self.start, self.end, self.order = start, end, order
        self.general = general

    @property
    def n(self) -> int:
        return int(self.start.shape[0])

    def rec_idx(self) -> np.ndarray:
        if not self.rec_counts:
            return np.zeros(0, dtype=np.int64)
        ids = np.array([r for r, _ in self.rec_counts], dtype=np.int64)
        cnt = np.array([c for _, c in self.rec_counts], dtype=np.int64)
        return np.repeat(ids, cnt)

    def as_int64(self):
        return (self.rec_idx(), self.start.astype(np.int64), self.end.astype(np.int64),
                self.order.astype(np.int64))


def build_index(dap_host, records: Optional[Sequence[Tuple[str, int]]], order: bool,
                chunk_bytes: int = DEFAULT_CHUNK_BYTES, device=None, pos_first: int = 0,
                stats: Optional[dict] = None, segs: Optional[List[api.Segment]] = None,
                raw: bool = False, **tuning):
    """DAP on the host (int32 [L, C], row i = global position pos_first + i) ->
    index rows in the reference's print order (src/dap_to_bed.py --mem --overlap
    [--order]).  Returns (rec_idx, start, end, order) int64 numpy arrays, or an
    IndexRowsHost (int32 columns, no widening) with raw=True.

    `segs` (explicit record runs over the rows of dap_host, e.g. a position shard
    with halo rows) overrides the records/pos_first layout.  If the device reports
    the input as irregular (not valid matching statistics) the exact three-pass
    build is run on the whole DAP instead.
    """
    host = _as_host_tensor(dap_host)
    if host.dtype != torch.int32 or host.dim() != 2:
        raise MemoError("dap must be int32 [L, C]")
    L, C = host.shape
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if L == 0:
        z = np.zeros(0, dtype=np.int64)
        res = IndexRowsHost([], z.astype(np.int32), z.astype(np.uint32), z.astype(np.int32), False)
        return res if raw else res.as_int64()
    # explicit runs (position shards with halos) or the whole-file layout of index.sh:83
    segs_all = list(segs) if segs is not None else \
        api.segments_for_rows(records, pos_first, L)               # raises like the reference
    if not host.is_pinned():
        # pageable memory would serialise the copies with the kernels
        staged = _pinned("dap", host.numel() * 4)[:host.numel() * 4].view(torch.int32).view(L, C)
        staged.copy_(host)
        host = staged

    builder = api.IndexBuilder(dev)
    chunk_rows = max(1, min(L, chunk_bytes // (4 * C)))
    bounds = list(range(0, L, chunk_rows)) + [L]
    n_chunks = len(bounds) - 1
    ring = [torch.empty((chunk_rows + 1, C), dtype=torch.int32, device=dev) for _ in range(min(RING, n_chunks))]
    cap = max(1024, int(chunk_rows * C * 0.02) + 2 * C)
    outs = torch.empty((n_chunks, 3, cap), dtype=torch.int32, device=dev)
    results = torch.zeros((n_chunks, RES_SLOTS), dtype=torch.int64, device=dev)
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    copy_stream.wait_stream(main)
    done = [None] * n_chunks
    chunk_segs, seg_ends = [], []

    def enqueue(i, cap_i, out_i):
        a, b = bounds[i], bounds[i + 1]
        lo = a - 1 if a > 0 else a                   # the halo row travels with the chunk
        buf = ring[i % len(ring)]
        if i >= len(ring):
            copy_stream.wait_event(done[i - len(ring)])
        with torch.cuda.stream(copy_stream):
            buf[:b - lo].copy_(host[lo:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        main.wait_event(ev)
        cs = _clip_segments(segs_all, a, b, a - lo)
        soe = torch.empty(max(len(cs), 1), dtype=torch.int64, device=dev)
        builder.launch(buf, C, cs, order, (out_i[0], out_i[1], out_i[2]), soe,
                       result=results[i], **tuning)
        done[i] = torch.cuda.Event()
        done[i].record(main)
        return cs, soe

    for i in range(n_chunks):
        cs, soe = enqueue(i, cap, outs[i])
        chunk_segs.append(cs)
        seg_ends.append(soe)
    res_host = results.cpu()                        # synchronises
    if bool((res_host[:, RES_IRREGULAR] != 0).any()):
        out = _build_index_general(host, C, segs_all, order, dev, builder, tuning)
        if stats is not None:
            stats.update(general=True, n_out=out.n)
        return out if raw else out.as_int64()
    counts = res_host[:, RES_N_OUT].tolist()
    big = {}
    for i, n in enumerate(counts):                  # denser than expected: redo with exact room
        if n > cap:
            big[i] = torch.empty((3, n), dtype=torch.int32, device=dev)
            copy_stream.wait_stream(main)          # the ring slot may still be in use by a retry
            chunk_segs[i], seg_ends[i] = enqueue(i, n, big[i])
    total = int(sum(counts))
    arr = None
    if total:
        blk, raw_bytes = _pinned_result(12 * total)
        stage = blk.view(torch.int32).view(3, total)
        arr = raw_bytes.view(np.int32).reshape(3, total)
        off = 0
        for i, n in enumerate(counts):
            if n:
                src = big[i] if i in big else outs[i]
                stage[:, off:off + n].copy_(src[:, :n], non_blocking=True)
                off += n
    width = max(len(cs) for cs in chunk_segs)
    seg_end_host = torch.stack([torch.nn.functional.pad(s, (0, width - s.numel())) for s in seg_ends]).cpu() \
        if width else None                           # one copy; synchronises the row copies too
    torch.cuda.current_stream(dev).synchronize()
    rec_counts = []
    for ci, cs in enumerate(chunk_segs):
        prev = 0
        for s, e in zip(cs, seg_end_host[ci].tolist()):
            if e > prev:
                if rec_counts and rec_counts[-1][0] == s.rec_id:
                    rec_counts[-1] = (s.rec_id, rec_counts[-1][1] + e - prev)
                else:
                    rec_counts.append((s.rec_id, e - prev))
            prev = e
    if total:
        out = IndexRowsHost(rec_counts, arr[0], arr[1].view(np.uint32), arr[2], False)
    else:
        z = np.zeros(0, dtype=np.int32)
        out = IndexRowsHost([], z, z.view(np.uint32), z.copy(), False)
    if stats is not None:
        stats.update(general=False, n_out=total, chunks=n_chunks)
    return out if raw else out.as_int64()


def _build_index_general(host, C, segs_all, order, dev, builder, tuning) -> IndexRowsHost:
    """Irregular input: the whole DAP goes to the device and the exact build runs."""
    dap = torch.empty(tuple(host.shape), dtype=torch.int32, device=dev)
    dap.copy_(host, non_blocking=True)
    res = builder.build(dap, C, segs_all, order, force_general=True, **tuning)
    n = res.n
    counts = np.diff(np.concatenate([[0], res.seg_out_end.numpy()]))
    rec_counts = [(rid, int(c)) for rid, c in zip(res.seg_rec_id, counts) if c]
    return IndexRowsHost(rec_counts, res.start[:n].cpu().numpy(),
                         res.end[:n].cpu().numpy().view(np.uint32), res.order[:n].cpu().numpy(), True)


def _rows_to_device(f1, f2, f3, dev, trusted=False):
    f1 = np.asarray(f1); f2 = np.asarray(f2); f3 = np.asarray(f3)
    if not trusted:
        if f1.size and (f1.min() < 0 or f1.max() > 2**31 - 1 or f2.min() < 0 or f2.max() > 2**32 - 1):
            raise MemoError("index row coordinates out of range")
        if f1.size and (f2.astype(np.int64) < f1.astype(np.int64)).any():
            raise MemoError("index rows with f2 < f1 are not MEMO index rows")
        if f1.size > 1 and (np.diff(f1.astype(np.int64)) < 0).any():
            o = np.argsort(f1, kind="stable")                # painting is order independent
            f1, f2, f3 = f1[o], f2[o], f3[o]
        f3 = np.clip(f3, -1, 2**31 - 1)                      # out-of-range ids are rejected on device
    n = f1.size
    if n == 0:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return z, z.clone(), z.clone()
    cols = []
    for a in (f1, f2, f3):
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        t = torch.from_numpy(a) if (a.dtype == np.int32 and a.flags.c_contiguous) else None
        cols.append(t if t is not None and t.is_pinned() else None)
    if all(c is not None for c in cols):              # rows straight from build_index: already pinned
        return tuple(c.to(dev, non_blocking=True) for c in cols)
    stage = _pinned("qrows", 12 * n)[:12 * n].view(torch.int32).view(3, n)
    view = stage.numpy()
    view[0] = f1                                      # same-kind casts into the pinned block
    view[1] = f2.astype(np.uint32, copy=False).view(np.int32) if f2.dtype != np.int32 else f2
    view[2] = f3
    d = stage.to(dev, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()      # the staging block is reused by the next call
    return d[0], d[1], d[2]


def query(f1, f2, f3, q_start: int, q_end: int, k: int, n_docs: int, membership: bool,
          device=None, as_text: bool = False, raw: bool = False, trusted: bool = False):
    """k-mer query over [q_start, q_end) from one record's index rows given as
    host arrays (src/memo_query.py:42-71).  Returns the conservation vector
    (int64 numpy; uint8/uint16 with raw=True) or the membership matrix (uint8
    [W, n_docs]); with as_text the exact bytes the reference writes to its output
    file.  trusted=True skips the host-side validation of rows that come straight
    from build_index."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if q_end < q_start:
        raise ValueError("negative dimensions are not allowed")      # np.ones/zeros in memo_init
    t1, t2, t3 = _rows_to_device(f1, f2, f3, dev, trusted)
    if as_text and q_end == q_start:
        return b"" if membership else b"\n"       # print(*[], sep='\n') still writes a newline
    if membership:
        bits = api.query_membership(t1, t2, t3, q_start, q_end, k, n_docs)
        if as_text:
            return api.format_membership(bits, n_docs)
        return api.unpack_membership(bits.cpu().numpy(), n_docs)
    out = api.query_conservation(t1, t2, t3, q_start, q_end, k, n_docs)
    if as_text:
        return api.format_conservation(out)
    nbytes = out.numel() * out.element_size()
    blk, raw_bytes = _pinned_result(nbytes)
    blk.view(out.dtype).copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    host = raw_bytes.view(np.uint16 if out.dtype == torch.int16 else np.uint8)
    return host if raw else host.astype(np.int64)


def query_sweep(f1, f2, f3, q_start: int, q_end: int, ks: Sequence[int], n_docs: int, membership: bool,
                device=None, trusted: bool = False) -> dict:
    """The same window queried for several k (BASELINE configs[4]: k = 15 .. 101): the index
    rows go to the device once, every k is one query launch over them.  Returns {k: result}
    with the results of `query` (conservation: uint8/uint16 numpy vectors; membership: uint8
    [W, n_docs] matrices)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if q_end < q_start:
        raise ValueError("negative dimensions are not allowed")
    t1, t2, t3 = _rows_to_device(f1, f2, f3, dev, trusted)
    W = q_end - q_start
    ws = torch.empty(max(api._lib.load().memo_query_workspace_bytes(W), 1), dtype=torch.uint8, device=dev)
    res = {}
    for k in ks:
        if membership:
            bits = api.query_membership(t1, t2, t3, q_start, q_end, k, n_docs, workspace=ws)
            res[k] = api.unpack_membership(bits.cpu().numpy(), n_docs)
        else:
            out = api.query_conservation(t1, t2, t3, q_start, q_end, k, n_docs, workspace=ws).cpu().numpy()
            res[k] = out.view(np.uint16) if out.dtype == np.int16 else out
    return res
