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


class IndexStream:
    """Streaming DAP -> index rows with O(chunk) memory, the shape of the reference's row loop
    (src/dap_to_bed.py:14-18,116-134: one pass, rows in, BED rows out).

    feed() takes blocks of consecutive DAP rows (int32 [n, C], any sizes); they are packed into
    a ring of pinned chunks (each carries the row before it as its halo), copied to the device
    on a copy stream while the previous chunk builds, and the index rows of every chunk come
    back through `on_rows(rec_counts, start, end, order)` in the reference's print order
    (rec_counts = [(record index, rows)], start / end / order = int32 / uint32 / int32 numpy
    views that are only valid during the call).  A chunk is submitted when the row after it
    arrives (or at finish()): whether its last run gets chr-end rows depends on that.

    Input that is not valid matching statistics: from the first chunk the device flags, the
    stream switches to the exact three-pass build, chunk by chunk, with the per-column carry
    handed from chunk to chunk (everything before that chunk was regular, for which the single-
    pass rows are exact and the carry is the sorted MEM ends of the halo row)."""

    def __init__(self, records: Optional[Sequence[Tuple[str, int]]], order: bool, n_cols: int, on_rows,
                 device=None, chunk_bytes: int = DEFAULT_CHUNK_BYTES, pos_first: int = 0,
                 segs_all: Optional[List[api.Segment]] = None, on_device: bool = False,
                 first_halo=None, **tuning):
        """segs_all: explicit record runs over the rows fed (row 0 = the first row fed), e.g. a
        position shard with halo rows, instead of the records / pos_first layout of index.sh:83.
        on_device: on_rows receives a DEVICE tensor int32 [3, n] (valid until the stream's main
        CUDA stream has run RING - 1 more chunks) instead of host views.
        first_halo: the DAP row before the first row fed (int [C]) when the stream is a position
        shard that starts inside a record: the first run then continues that record."""
        self.records, self.order, self.C, self.on_rows, self.tuning = list(records or []), order, n_cols, on_rows, tuning
        self.segs_all, self.on_device = segs_all, on_device
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.builder = api.IndexBuilder(self.dev)
        C = n_cols
        self.chunk_rows = max(1, chunk_bytes // (4 * C))
        self.cap = max(1024, int(self.chunk_rows * C * 0.025) + 2 * C)
        self.main = torch.cuda.current_stream(self.dev)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.back_stream = torch.cuda.Stream(self.dev)
        self.slots = []                     # allocated on demand (a short input needs one)
        self.k = 0                          # chunks submitted so far
        self.collected = 0                  # chunks handed to on_rows so far
        self.fill = 0                       # rows waiting in the current slot
        self.pos = pos_first                # global position of the next row to arrive
        self.pos_stream0 = pos_first
        self.starts, acc = set(), 0
        for _, length in self.records:
            self.starts.add(acc)
            acc += length
        self.on_dev = False                 # the current chunk is being filled device to device (feed_device)
        self.prev_last_row_dev = None       # last row of the previous chunk, on the device
        self.src = None                     # rows of the current chunk taken in place (pinned caller memory)
        self.last_row = None                # last row fed so far
        # last row of the chunk submitted before the current one
        self.prev_last_row = None if first_halo is None else np.asarray(first_halo, dtype=np.int32).copy()
        self.first_halo = first_halo is not None
        self.general = False                # exact build from here on (irregular input seen)
        self.carry = None                   # device int64 [C]: per-column end of the last flagged MEM
        self.n_out = 0
        self.n_chunks_general = 0

    # ---- buffers
    def _slot(self, i):
        while len(self.slots) <= i % RING:
            C, n = self.C, self.chunk_rows + 1
            self.slots.append({
                "pin": None,                # pinned staging of the chunk's rows (allocated when a block needs it)
                "pin_halo": torch.empty((1, C), dtype=torch.int32, pin_memory=True),
                "dev": torch.empty((n, C), dtype=torch.int32, device=self.dev),      # row 0 = halo row
                "out": torch.empty((3, self.cap), dtype=torch.int32, device=self.dev),
                "pin_out": None if self.on_device else torch.empty((3, self.cap), dtype=torch.int32, pin_memory=True),
                "res": torch.zeros(RES_SLOTS, dtype=torch.int64, device=self.dev),
                "pin_res": torch.zeros(RES_SLOTS, dtype=torch.int64, pin_memory=True),
            })
        return self.slots[i % RING]

    def _fill_slot(self, i):
        """Slot of chunk i, its pinned input free to be written (the copy of the chunk that used
        the slot before has left it)."""
        slot = self._slot(i)
        ev = slot.pop("h2d", None)
        if ev is not None:
            ev.synchronize()
        return slot

    def feed_device(self, block: torch.Tensor) -> None:
        """Rows [pos, pos + n) of the DAP already on the device (int32 [n, C], e.g. from the
        device text parser): copied into the chunk buffers on the stream's CUDA stream; the block
        may be overwritten as soon as this returns (stream order)."""
        if block.dim() != 2 or block.shape[1] != self.C or block.dtype != torch.int32 or not block.is_cuda:
            raise MemoError(f"device DAP block must be int32 [n, {self.C}] on the GPU")
        done = 0
        while done < block.shape[0]:
            if self.fill == self.chunk_rows:
                self._submit(last=False)
            slot = self._slot(self.k)
            if self.fill == 0:
                pos0 = self.pos
                if self.segs_all is not None:
                    slot["base"] = 1 if pos0 > self.pos_stream0 else 0
                else:
                    slot["base"] = 1 if (pos0 > self.pos_stream0 or self.first_halo) and pos0 not in self.starts else 0
                self.on_dev = True
            n = min(block.shape[0] - done, self.chunk_rows - self.fill)
            b = slot["base"] + self.fill
            slot["dev"][b:b + n].copy_(block[done:done + n], non_blocking=True)
            self.fill += n
            self.pos += n
            done += n

    def feed(self, block) -> None:
        """Rows [pos, pos + n) of the DAP.  A pinned torch tensor is copied to the device straight
        from where it is (it has to stay alive until finish()); a CUDA tensor is taken on the
        device (feed_device); anything else goes through the stream's pinned staging."""
        if isinstance(block, torch.Tensor) and block.is_cuda:
            return self.feed_device(block)
        direct = isinstance(block, torch.Tensor) and block.is_pinned() and block.is_contiguous()
        if isinstance(block, torch.Tensor) and not direct:
            block = block.numpy()
        if not direct:
            block = np.asarray(block)
        if block.ndim != 2 or block.shape[1] != self.C:
            raise MemoError(f"DAP block must be [n, {self.C}]")
        if block.shape[0] == 0:
            return
        done = 0
        while done < block.shape[0]:
            if self.fill == self.chunk_rows:
                self._submit(last=False)
            n = min(block.shape[0] - done, self.chunk_rows - self.fill)
            if direct and self.fill == 0 and n == min(self.chunk_rows, block.shape[0] - done):
                self.src = block[done:done + n]            # a whole chunk (or the block's tail) in place
            else:
                slot = self._fill_slot(self.k)
                if slot["pin"] is None:
                    slot["pin"] = torch.empty((self.chunk_rows, self.C), dtype=torch.int32, pin_memory=True)
                if self.src is not None:                   # rows taken in place earlier join the staging
                    m = self.src.shape[0]
                    slot["pin"].numpy()[:m] = self.src.numpy()
                    self.src = None
                piece = block[done:done + n]
                slot["pin"].numpy()[self.fill:self.fill + n] = piece.numpy() if direct else piece
            self.fill += n
            self.pos += n
            done += n
            tail = block[done - 1]
            self.last_row = (tail.numpy() if direct else np.asarray(tail)).copy()

    def finish(self, final: bool = True) -> int:
        """Flushes the last chunk and every pending result; returns the number of index rows.
        final=False: the DAP goes on in another position shard (no chr-end rows after a last run
        that stops inside its record)."""
        if self.fill:
            self._submit(last=final)
        while self.collected < self.k:
            self._collect()
        return self.n_out

    # ---- one chunk
    def _segments(self, pos0, n, halo, last):
        segs = api.segments_for_rows(self.records, pos0, n, buffer_row0=1 if halo else 0,
                                     primed_first=not halo, chr_end_last=True)
        tail = segs[-1]
        if not last and tail.pos0 + tail.n_rows < tail.rec_len:
            tail.flags &= ~api.MEMO_SEG_CHR_END          # the record goes on in the next chunk
        return segs

    def _submit(self, last: bool) -> None:
        while self.k - self.collected >= RING - 1:        # (one slot is being filled, RING - 1 in flight)
            self._collect()
        slot = self._fill_slot(self.k)
        n = self.fill
        pos0 = self.pos - n
        if self.segs_all is not None:
            # explicit runs over the rows fed: the previous row always travels with the chunk
            a = pos0 - self.pos_stream0
            halo = a > 0
            segs = _clip_segments(self.segs_all, a, a + n, 1 if halo else 0)
        else:
            # the halo row = the record's previous row
            halo = (pos0 > self.pos_stream0 or self.first_halo) and pos0 not in self.starts
            segs = self._segments(pos0, n, halo, last)
        slot.update(n=n, pos0=pos0, halo=halo, last=last, segs=segs)
        if self.on_dev:
            # rows were copied device to device as they arrived (feed_device); the halo row too
            assert slot["base"] == (1 if halo else 0)
            if halo:
                prev = self.prev_last_row_dev
                if prev is None:                           # (the shard's first halo row came from the host)
                    prev = torch.from_numpy(self.prev_last_row.reshape(1, -1)).to(self.dev)
                slot["dev"][0:1].copy_(prev, non_blocking=True)
            # next chunk's halo = this chunk's last row
            self.prev_last_row_dev = slot["dev"][slot["base"] + n - 1:slot["base"] + n].clone()
            self.on_dev = False
        else:
            src = self.src if self.src is not None else slot["pin"][:n]
            self.src = None
            if halo:
                slot["pin_halo"].numpy()[0] = self.prev_last_row
            with torch.cuda.stream(self.copy_stream):
                # (the build that read this slot's device buffer last is RING chunks back: collected)
                # device chunk = [halo row,] rows: always starts at the (16-byte aligned) buffer base
                if halo:
                    slot["dev"][0:1].copy_(slot["pin_halo"], non_blocking=True)
                base = 1 if halo else 0
                slot["dev"][base:base + n].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            slot["h2d"] = ev
            self.main.wait_event(ev)
            self.prev_last_row = self.last_row             # next chunk's halo = this chunk's last row
        if not self.general and segs:
            self._launch_fast(slot)
        slot["general"] = self.general
        self.k += 1
        self.fill = 0
        if self.general:
            while self.collected < self.k:                 # the exact build runs chunk by chunk
                self._collect()

    def _launch_fast(self, slot) -> None:
        n_buf = slot["n"] + (1 if slot["halo"] else 0)
        soe = torch.empty(max(len(slot["segs"]), 1), dtype=torch.int64, device=self.dev)
        self.builder.launch(slot["dev"][:n_buf], self.C, slot["segs"], self.order,
                            (slot["out"][0], slot["out"][1], slot["out"][2]), soe, result=slot["res"], **self.tuning)
        ev = torch.cuda.Event()
        ev.record(self.main)
        with torch.cuda.stream(self.back_stream):
            self.back_stream.wait_event(ev)
            slot["pin_res"].copy_(slot["res"], non_blocking=True)
            # (into PINNED memory: `.to("cpu", non_blocking=True)` lands in pageable memory, and a
            #  device -> pageable copy returns only when it is done -- one host stall per chunk)
            pin = slot.get("pin_soe_buf")
            if pin is None or pin.numel() < soe.numel():
                pin = slot["pin_soe_buf"] = torch.empty(max(64, soe.numel()), dtype=torch.int64, pin_memory=True)
            slot["pin_soe"] = pin[:soe.numel()]
            slot["pin_soe"].copy_(soe, non_blocking=True)
            if not self.on_device:
                slot["pin_out"].copy_(slot["out"], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.back_stream)
        slot["soe"], slot["done"] = soe, done

    def _collect(self) -> None:
        slot = self._slot(self.collected)
        self.collected += 1
        if not slot["segs"]:
            return                                         # rows that belong to no run
        if not slot["general"]:
            slot["done"].synchronize()
            n_out, irregular = int(slot["pin_res"][RES_N_OUT]), bool(slot["pin_res"][RES_IRREGULAR])
            if irregular:
                # exact build from this chunk on; chunks already submitted behind it are redone too
                self.general = True
                pending = [self._slot(i) for i in range(self.collected, self.k)]
                self._run_general(slot)
                for other in pending:
                    other["general"] = True
                return
            if n_out > self.cap:                          # denser than expected: once more with exact room
                big = torch.empty((3, n_out), dtype=torch.int32, device=self.dev)
                self.builder.launch(slot["dev"][:slot["n"] + (1 if slot["halo"] else 0)], self.C, slot["segs"], self.order,
                                    (big[0], big[1], big[2]), slot["soe"], **self.tuning)
                rows = big if self.on_device else big.cpu().numpy()
                seg_ends = slot["soe"].cpu().tolist()
            else:
                rows = slot["out"][:, :n_out] if self.on_device else slot["pin_out"].numpy()[:, :n_out]
                seg_ends = slot["pin_soe"].tolist()
            self._deliver(slot, rows, seg_ends)
        else:
            self._run_general(slot)

    def _run_general(self, slot) -> None:
        C = self.C
        buf = slot["dev"][:slot["n"] + (1 if slot["halo"] else 0)]
        segs = slot["segs"]
        soe = torch.empty(max(len(segs), 1), dtype=torch.int64, device=self.dev)
        cin = None
        if slot["halo"] and not (segs[0].flags & api.MEMO_SEG_PRIMED):
            if self.carry is None:
                # everything before this chunk was regular: the last flagged MEM of a column ends
                # where the (sorted) MEM ends of the run's previous row do
                e = buf[segs[0].row_begin - 1].to(torch.int64) + (segs[0].pos0 - 1)
                self.carry = torch.sort(e, descending=True).values if self.order else e
            cin = self.carry.to(torch.int32)              # wraps to the uint32 bit pattern
        cout = torch.full((C,), -1, dtype=torch.int32, device=self.dev)
        self.builder.launch(buf, C, segs, self.order, None, soe, general=True, carry_in=cin, carry_out=cout,
                            **self.tuning)
        n_out, _, _ = self.builder.result()
        out = torch.empty((3, max(n_out, 1)), dtype=torch.int32, device=self.dev)
        self.builder.launch(buf, C, segs, self.order, (out[0], out[1], out[2]), soe, general=True, carry_in=cin,
                            carry_out=cout, **self.tuning)
        self.builder.result()
        co = cout.to(torch.int64) & 0xFFFFFFFF
        keep = self.carry if (self.carry is not None and cin is not None) else torch.full_like(co, 0xFFFFFFFF)
        self.carry = torch.where(co != 0xFFFFFFFF, co, keep)
        self.n_chunks_general += 1
        self._deliver(slot, out[:, :n_out] if self.on_device else out[:, :n_out].cpu().numpy(), soe.cpu().tolist())

    def _deliver(self, slot, rows, seg_ends) -> None:
        rec_counts, prev = [], 0
        for s, e in zip(slot["segs"], seg_ends):
            if e > prev:
                if rec_counts and rec_counts[-1][0] == s.rec_id:
                    rec_counts[-1] = (s.rec_id, rec_counts[-1][1] + e - prev)
                else:
                    rec_counts.append((s.rec_id, e - prev))
            prev = e
        self.n_out += prev
        if prev and self.on_device:
            self.on_rows(rec_counts, rows[:, :prev])
        elif prev:
            self.on_rows(rec_counts, rows[0, :prev], rows[1, :prev].view(np.uint32), rows[2, :prev])


def build_index_streaming(blocks, records, order: bool, on_rows, n_cols: Optional[int] = None,
                          pos_first: int = 0, device=None, chunk_bytes: int = DEFAULT_CHUNK_BYTES,
                          stats: Optional[dict] = None, first_halo=None, final: bool = True,
                          on_device: bool = False, **tuning) -> int:
    """Index rows of a DAP that arrives as an iterator of int32 [n, C] blocks of consecutive rows
    (first row = global position pos_first); on_rows / on_device as in IndexStream.  first_halo / final: the
    stream is one position shard of a larger DAP (IndexStream).  Returns the row count."""
    stream = None
    for block in blocks:
        if stream is None:
            C = block.shape[1] if n_cols is None else n_cols
            stream = IndexStream(records, order, C, on_rows, device=device, chunk_bytes=chunk_bytes,
                                 pos_first=pos_first, first_halo=first_halo, on_device=on_device, **tuning)
        stream.feed(block)
    if stream is None:
        return 0
    n = stream.finish(final)
    if stats is not None:
        stats.update(general=stream.general, n_out=n, chunks=stream.k, chunks_general=stream.n_chunks_general)
    return n


def build_index(dap_host, records: Optional[Sequence[Tuple[str, int]]], order: bool,
                chunk_bytes: int = DEFAULT_CHUNK_BYTES, device=None, pos_first: int = 0,
                stats: Optional[dict] = None, segs: Optional[List[api.Segment]] = None,
                raw: bool = False, **tuning):
    """DAP on the host (int32 [L, C], row i = global position pos_first + i) ->
    index rows in the reference's print order (src/dap_to_bed.py --mem --overlap
    [--order]).  Returns (rec_idx, start, end, order) int64 numpy arrays, or an
    IndexRowsHost (int32 columns in one pinned block, no widening) with raw=True.

    `segs` (explicit record runs over the rows of dap_host, e.g. a position shard
    with halo rows) overrides the records/pos_first layout.  The DAP streams through
    IndexStream: device memory and pinned staging are O(chunk_bytes), whatever L;
    input that is not valid matching statistics switches to the exact build on the way.
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
    if segs is None:
        api.segments_for_rows(records, pos_first, L)               # raises like the reference
    # index rows land in one pinned block, copied there chunk by chunk as they are built
    state = {"blk": None, "arr": None, "n": 0, "cap": 0, "rec_counts": []}

    def room(n_more):
        need = state["n"] + n_more
        if need <= state["cap"]:
            return
        cap = max(need + need // 2, 1 << 16)
        blk, raw_bytes = _pinned_result(12 * cap)
        if state["n"]:
            torch.cuda.current_stream(dev).synchronize()
            raw_bytes.view(np.int32).reshape(3, cap)[:, :state["n"]] = state["arr"][:, :state["n"]]
        state.update(blk=blk.view(torch.int32).view(3, cap), arr=raw_bytes.view(np.int32).reshape(3, cap), cap=cap)

    def on_rows(rec_counts, dev_rows):
        n = dev_rows.shape[1]
        room(n)
        # (one contiguous copy per column: a strided [3, n] copy across devices goes through
        #  temporaries in pageable memory and blocks the host)
        for r in range(3):
            state["blk"][r, state["n"]:state["n"] + n].copy_(dev_rows[r], non_blocking=True)
        state["n"] += n
        for rid, c in rec_counts:
            if state["rec_counts"] and state["rec_counts"][-1][0] == rid:
                state["rec_counts"][-1] = (rid, state["rec_counts"][-1][1] + c)
            else:
                state["rec_counts"].append((rid, c))

    chunk_rows = max(1, chunk_bytes // (4 * C))
    room(int(L * C * 0.022) + 4 * C)
    stream = IndexStream(records, order, C, on_rows, device=dev, chunk_bytes=chunk_bytes, pos_first=pos_first,
                         segs_all=list(segs) if segs is not None else None, on_device=True, **tuning)
    for a in range(0, L, chunk_rows):
        stream.feed(host[a:a + chunk_rows])
    total = stream.finish()
    torch.cuda.current_stream(dev).synchronize()
    if total:
        arr = state["arr"]
        out = IndexRowsHost(state["rec_counts"], arr[0, :total], arr[1, :total].view(np.uint32), arr[2, :total],
                            stream.general)
    else:
        z = np.zeros(0, dtype=np.int32)
        out = IndexRowsHost([], z, z.view(np.uint32), z.copy(), stream.general)
    if stats is not None:
        stats.update(general=stream.general, n_out=total, chunks=stream.k)
    return out if raw else out.as_int64()


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
    n4 = (n + 3) // 4 * 4                             # columns 16-byte aligned on the device
    stage = _pinned("qrows", 12 * n4)[:12 * n4].view(torch.int32).view(3, n4)
    view = stage.numpy()
    view[0, :n] = f1                                  # same-kind casts into the pinned block
    view[1, :n] = f2.astype(np.uint32, copy=False).view(np.int32) if f2.dtype != np.int32 else f2
    view[2, :n] = f3
    d = stage.to(dev, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()      # the staging block is reused by the next call
    return d[0, :n], d[1, :n], d[2, :n]


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
    """The same window queried for several k (BASELINE configs[4]: k = 15 .. 101): the index rows
    go to the device once and all k are answered by ONE launch per 16 values (memo_query_sweep:
    tiles, row search and rows shared) and one copy back.  Returns {k: result} with the results
    of `query` (conservation: uint8 numpy vectors; membership: uint8 [W, n_docs] matrices).
    More than 255 genomes (uint16 results) take one launch per k."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if q_end < q_start:
        raise ValueError("negative dimensions are not allowed")
    t1, t2, t3 = _rows_to_device(f1, f2, f3, dev, trusted)
    W = q_end - q_start
    ks = list(ks)
    nw = (n_docs + 31) // 32
    fused = (n_docs <= 255 if not membership else (W * nw) % 4 == 0) and \
        q_end + max(ks, default=1) < 2**31 - 2**17 and all(t.data_ptr() % 16 == 0 for t in (t1, t2, t3))
    if fused:
        out = api.query_sweep(t1, t2, t3, q_start, q_end, ks, n_docs, membership).cpu().numpy()
        if membership:
            return {k: api.unpack_membership(out[i], n_docs) for i, k in enumerate(ks)}
        return {k: out[i, :W] for i, k in enumerate(ks)}
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
