"""Host-buffer API: the calls the drop-in scripts (and bench.py's e2e leg) make.

Inputs and outputs are HOST arrays; host<->device copies happen inside.  The
DAP is streamed to the device in row chunks on a copy stream while the build
kernel for the previous chunk runs on the compute stream.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import api
from ._lib import MemoError

DEFAULT_CHUNK_ROWS = 1 << 23


def _as_host_tensor(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.is_cuda:
        raise MemoError("host API expects host buffers")
    return t.contiguous()


def build_index(dap_host, records: Optional[Sequence[Tuple[str, int]]], order: bool,
                chunk_rows: int = DEFAULT_CHUNK_ROWS, device=None, pos_first: int = 0,
                return_device: bool = False, stats: Optional[dict] = None,
                segs: Optional[List[api.Segment]] = None, **tuning):
    """DAP on the host (int32 [L, C], row i = global position pos_first + i) ->
    index rows (rec_idx, start, end, order) as int64 numpy arrays in the
    reference's print order (src/dap_to_bed.py --mem --overlap [--order]).

    The whole DAP is kept on the device (it must fit); chunks are copied
    asynchronously and each chunk is built as soon as it has landed.  `segs`
    (explicit record runs, e.g. a position shard with halo rows) overrides the
    records/pos_first layout.  If the
    device reports the input as irregular (not valid matching statistics) the
    exact three-pass build is re-run on the resident buffer.
    """
    host = _as_host_tensor(dap_host)
    if host.dtype != torch.int32 or host.dim() != 2:
        raise MemoError("dap must be int32 [L, C]")
    L, C = host.shape
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    builder = api.IndexBuilder(dev)
    if L == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), z.copy(), z.copy()
    # explicit runs (position shards with halos) or the whole-file layout of index.sh:83
    segs_all = list(segs) if segs is not None else \
        api.segments_for_rows(records, pos_first, L)               # raises like the reference
    dap = torch.empty((L, C), dtype=torch.int32, device=dev)
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    bounds = list(range(0, L, chunk_rows)) + [L]
    events = []
    copy_stream.wait_stream(main)
    with torch.cuda.stream(copy_stream):
        for i in range(len(bounds) - 1):
            a, b = bounds[i], bounds[i + 1]
            dap[a:b].copy_(host[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            events.append(ev)

    parts = []        # (out tensors, seg_out_end, segs, n)
    general = False
    for i in range(len(bounds) - 1):
        a, b = bounds[i], bounds[i + 1]
        main.wait_event(events[i])
        segs = _clip_segments(segs_all, a, b)
        seg_out_end = torch.zeros(len(segs), dtype=torch.int64, device=dev)
        cap = max(1024, int((b - a) * C * 0.02) + C * (len(segs) + 1))
        while True:
            out = tuple(torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3))
            builder.launch(dap, C, segs, order, out, seg_out_end, **tuning)
            n, irregular, _ = builder.result()
            if irregular:
                general = True
                break
            if n > cap:
                cap = n
                continue
            break
        if general:
            break
        parts.append((out, seg_out_end.cpu(), segs, n))
    if general:
        main.wait_stream(copy_stream)
        res = builder.build(dap, C, segs_all, order, force_general=True, **tuning)
        if stats is not None:
            stats.update(general=True, n_out=res.n)
        return res if return_device else res.to_host()
    if stats is not None:
        stats.update(general=False, n_out=sum(p[3] for p in parts))
    if return_device:
        n = sum(p[3] for p in parts)
        cat = [torch.cat([p[0][j][:p[3]] for p in parts]) for j in range(3)]
        seg_ids, seg_end, acc = [], [], 0
        for out, soe, segs, cnt in parts:
            seg_ids += [s.rec_id for s in segs]
            seg_end += [acc + int(x) for x in soe.tolist()]
            acc += cnt
        return api.IndexRows(cat[0], cat[1], cat[2], torch.tensor(seg_end, dtype=torch.int64),
                             seg_ids, n, False, 0, False)
    rec, start, end, col = [], [], [], []
    for out, soe, segs, n in parts:
        start.append(out[0][:n].cpu().numpy().astype(np.int64))
        end.append(out[1][:n].cpu().numpy().view(np.uint32).astype(np.int64))
        col.append(out[2][:n].cpu().numpy().astype(np.int64))
        counts = np.diff(np.concatenate([[0], soe.numpy()]))
        rec.append(np.repeat(np.array([s.rec_id for s in segs], dtype=np.int64), counts))
    return (np.concatenate(rec), np.concatenate(start), np.concatenate(end), np.concatenate(col))


def _clip_segments(segs_all: List[api.Segment], a: int, b: int) -> List[api.Segment]:
    """Record runs restricted to buffer rows [a, b): a run cut on the left
    continues its record (halo = row a-1, already resident), a run cut on the
    right does not emit its chr-end rows yet."""
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
        out.append(api.Segment(row_begin=lo, n_rows=hi - lo, pos0=s.pos0 + (lo - s.row_begin),
                               rec_len=s.rec_len, rec_id=s.rec_id, flags=flags))
    return out


def _rows_to_device(f1, f2, f3, dev):
    f1 = np.asarray(f1); f2 = np.asarray(f2); f3 = np.asarray(f3)
    if f1.size and (f1.min() < 0 or f1.max() > 2**31 - 1 or f2.min() < 0 or f2.max() > 2**32 - 1):
        raise MemoError("index row coordinates out of range")
    if f1.size and (f2 < f1).any():
        raise MemoError("index rows with f2 < f1 are not MEMO index rows")
    if f1.size > 1 and (np.diff(f1) < 0).any():
        o = np.argsort(f1, kind="stable")                # painting is order independent
        f1, f2, f3 = f1[o], f2[o], f3[o]
    f3c = np.clip(f3, -1, 2**31 - 1)                      # out-of-range ids are rejected on device
    t1 = torch.from_numpy(f1.astype(np.int32)).to(dev, non_blocking=True)
    t2 = torch.from_numpy(f2.astype(np.uint32).view(np.int32)).to(dev, non_blocking=True)
    t3 = torch.from_numpy(f3c.astype(np.int32)).to(dev, non_blocking=True)
    return t1, t2, t3


def query(f1, f2, f3, q_start: int, q_end: int, k: int, n_docs: int, membership: bool,
          device=None, as_text: bool = False):
    """k-mer query over [q_start, q_end) from one record's index rows given as
    host arrays (src/memo_query.py:42-71).  Returns the conservation vector
    (int64 numpy) or the membership matrix (uint8 [W, n_docs]); with as_text the
    exact bytes the reference writes to its output file."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if q_end < q_start:
        raise ValueError("negative dimensions are not allowed")      # np.ones/zeros in memo_init
    t1, t2, t3 = _rows_to_device(f1, f2, f3, dev)
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
    host = out.cpu().numpy()
    if host.dtype == np.int16:
        host = host.view(np.uint16)
    return host.astype(np.int64)
