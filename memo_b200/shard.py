"""Position-range sharding of the pivot across the GPUs of one box.

Shards are independent: rank r builds the index rows of its own position range
from its own DAP rows plus a one-row left halo (the flag compares a row with its
predecessor) and, for queries, a right halo of k_max-1 rows.  The only exchange
is an all-gather of one int64 per rank -- the number of index rows a rank owns --
whose exclusive prefix is the rank's offset in the ordered index.

Input that is not valid matching statistics (the reference never validates,
SURVEY A.1) needs one more exchange: the per-column end of the last flagged MEM
has to cross shard boundaries, so the ranks agree on the irregular verdict
(all-reduce of one flag), all-gather their per-column carries (n_cols x uint32
per rank) and re-run the exact three-pass build with the carry handed in.

Every collective here works on the tensors' own device: NCCL for CUDA tensors on
the GPU box, gloo for the CPU tensors of the world_size-2 tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import MEMO_SEG_CHR_END, MEMO_SEG_PRIMED, MemoError, Segment

NONE32 = 0xFFFFFFFF          # "no flagged row" in a carry vector (uint32 bits)


@dataclass
class ShardPlan:
    pos_lo: int            # owned global positions [pos_lo, pos_hi)
    pos_hi: int
    buf_lo: int            # global positions held in the rank's DAP buffer [buf_lo, buf_hi)
    buf_hi: int
    segs: List[Segment]    # record runs over the buffer (owned runs first, then halo runs)
    n_owned: int           # number of leading runs whose index rows this rank owns


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Near-equal contiguous split of [0, total)."""
    return (total * rank) // world, (total * (rank + 1)) // world


def _runs(records, lo, hi, buf_lo, first_continues):
    segs, acc = [], 0
    for rid, (_, length) in enumerate(records):
        a, b = max(lo, acc), min(hi, acc + length)
        if a < b:
            flags = 0
            if a == acc or (a == lo and not first_continues):
                flags |= MEMO_SEG_PRIMED
            if b == acc + length:
                flags |= MEMO_SEG_CHR_END
            segs.append(Segment(row_begin=a - buf_lo, n_rows=b - a, pos0=a - acc, rec_len=length,
                                rec_id=rid, flags=flags))
        acc += length
    return segs


def plan_shard(records: Sequence[Tuple[str, int]], total_rows: int, world: int, rank: int,
               halo_right: int = 0) -> ShardPlan:
    """Runs for rank `rank` of `world` over a DAP of `total_rows` rows (global
    positions 0..total_rows-1 over the concatenated records)."""
    lo, hi = shard_range(total_rows, world, rank)
    # a shard that starts inside a record needs that record's previous row
    starts, acc = set(), 0
    for _, length in records:
        starts.add(acc)
        acc += length
    needs_halo = lo > 0 and lo not in starts
    buf_lo = lo - 1 if needs_halo else lo
    buf_hi = min(hi + halo_right, total_rows)
    owned = _runs(records, lo, hi, buf_lo, first_continues=needs_halo)
    # a run that continues on the next rank leaves its chr-end rows to that rank; the run that
    # ends the DAP always emits them, also when the DAP stops inside the record
    # (src/dap_to_bed.py:133-134 runs after the last row whatever the record's .fai length)
    if owned and hi < total_rows:
        last = owned[-1]
        rec_end = sum(l for _, l in records[:last.rec_id + 1])
        if hi < rec_end:
            last.flags &= ~MEMO_SEG_CHR_END
    elif owned:
        owned[-1].flags |= MEMO_SEG_CHR_END
    halo = _runs(records, hi, buf_hi, buf_lo, first_continues=True) if buf_hi > hi else []
    for s in halo:
        s.flags &= ~MEMO_SEG_CHR_END            # halo rows only feed this rank's queries
    return ShardPlan(lo, hi, buf_lo, buf_hi, owned + halo, len(owned))


def ordered_offsets(n_local: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of the per-rank owned row counts (int64 tensor with one element,
    on the device of the backend).  Returns (counts[world], exclusive offset of
    this rank)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return n_local.reshape(1), torch.zeros((), dtype=torch.int64, device=n_local.device)
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=n_local.device)
    dist.all_gather_into_tensor(counts, n_local.reshape(1), group=group)
    rank = dist.get_rank(group)
    return counts, counts[:rank].sum()


def ordered_file_write(part_path: str, out_path: str, device, group=None) -> int:
    """The final ordered write of a sharded build: every rank holds its share of the output
    (BED text, in the reference's print order) in `part_path`; the byte counts are all-gathered,
    their exclusive prefix is a rank's offset in `out_path`, every rank copies its part there and
    removes it.  Returns the total size."""
    import os
    import shutil
    size = torch.tensor([os.path.getsize(part_path)], dtype=torch.int64, device=device)
    counts, offset = ordered_offsets(size, group)
    total = int(counts.sum().item())
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    if rank == 0:
        with open(out_path, "wb") as fh:
            fh.truncate(total)
    if multi:
        dist.barrier(group=group)
    with open(part_path, "rb") as src, open(out_path, "r+b") as dst:
        dst.seek(int(offset.item()))
        shutil.copyfileobj(src, dst, 16 << 20)
    if multi:
        dist.barrier(group=group)
    os.remove(part_path)
    return total


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def agree_irregular(irregular: bool, device, group=None) -> bool:
    """True on every rank iff any rank's shard is not valid matching statistics."""
    if _world(group) == 1:
        return bool(irregular)
    t = torch.tensor([1 if irregular else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def carry_in_from_gathered(carries: torch.Tensor, rank: int) -> torch.Tensor:
    """carries: int64 [world, C] of per-shard carry-outs (NONE32 = the shard flagged
    nothing in that column and handed nothing on).  The carry into `rank` is the
    closest earlier shard's value per column.  A shard hands on NONE32 only when
    it is a single run that continues its record, so the walk back never leaves
    the record of `rank`'s first run (a record's first row is always flagged)."""
    C = carries.shape[1]
    out = torch.full((C,), NONE32, dtype=torch.int64, device=carries.device)
    for q in range(rank):
        row = carries[q]
        out = torch.where(row != NONE32, row, out)
    return out


def exchange_carries(carry_out: torch.Tensor, group=None) -> torch.Tensor:
    """carry_out: int64 [C] (uint32 values) of this rank.  All-gather over the
    ranks; returns the carry into this rank (int64 [C], NONE32 where none)."""
    if _world(group) == 1:
        return torch.full_like(carry_out, NONE32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    allc = torch.empty((world, carry_out.numel()), dtype=carry_out.dtype, device=carry_out.device)
    dist.all_gather_into_tensor(allc, carry_out.reshape(1, -1).contiguous(), group=group)
    return carry_in_from_gathered(allc, rank)


def gather_index_rows(cols: torch.Tensor, counts: torch.Tensor, dst: int = 0, group=None):
    """The final ordered write: rank `dst` receives every rank's owned index rows
    (cols: int32 [3, n_local]; counts: the all-gathered row counts) and returns
    them concatenated in rank order = the reference's print order; the other ranks
    return None.  Point-to-point so that no rank pads to the largest shard."""
    if _world(group) == 1:
        return cols
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = [int(c) for c in counts.tolist()]
    if rank != dst:
        if n[rank]:
            dist.send(cols.contiguous(), dst=dst, group=group)
        return None
    out = torch.empty((3, sum(n)), dtype=cols.dtype, device=cols.device)
    off = 0
    for r in range(world):
        if n[r]:
            if r == rank:
                out[:, off:off + n[r]] = cols
            else:
                buf = torch.empty((3, n[r]), dtype=cols.dtype, device=cols.device)
                dist.recv(buf, src=r, group=group)
                out[:, off:off + n[r]] = buf
        off += n[r]
    return out


def build_index_sharded(dap: torch.Tensor, plan: ShardPlan, n_cols: int, order: bool,
                        builder=None, group=None, **tuning):
    """Index rows of this rank's position shard (`dap` = device rows
    [plan.buf_lo, plan.buf_hi) of the pivot), exact for every input.

    Returns (IndexRows, counts[world], offset): the rows this rank owns, every
    rank's row count and this rank's exclusive offset in the ordered index.
    """
    from . import api
    if not dap.is_cuda:
        raise MemoError("build_index_sharded needs the shard on a CUDA device (no CPU fallback)")
    dev = dap.device
    builder = api.IndexBuilder(dev) if builder is None else builder
    owned = plan.segs[:plan.n_owned]
    n_rows = plan.pos_hi - plan.pos_lo
    seg_out_end = torch.zeros(max(len(owned), 1), dtype=torch.int64, device=dev)
    cap = max(1024, int(n_rows * n_cols * 0.02) + n_cols * (len(owned) + 1))
    general = False
    while True:
        out = tuple(torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(3))
        builder.launch(dap, n_cols, owned, order, out, seg_out_end, **tuning)
        n, irregular, replays = builder.result()
        if n <= cap:
            break
        cap = n
    if agree_irregular(irregular, dev, group):
        general = True
        carry_out = torch.full((n_cols,), -1, dtype=torch.int32, device=dev)      # NONE32 bits
        builder.launch(dap, n_cols, owned, order, None, seg_out_end, general=True,
                       carry_out=carry_out, **tuning)
        n, _, _ = builder.result()
        carry64 = carry_out.to(torch.int64) & NONE32
        carry_in = exchange_carries(carry64, group)
        carry_in32 = carry_in.to(torch.int32)           # wraps back to the uint32 bit pattern
        out = tuple(torch.empty(max(n, 1), dtype=torch.int32, device=dev) for _ in range(3))
        while True:
            builder.launch(dap, n_cols, owned, order, out, seg_out_end, general=True,
                           carry_in=carry_in32, **tuning)
            n2, _, replays = builder.result()
            if n2 <= out[0].numel():
                n = n2
                break
            out = tuple(torch.empty(n2, dtype=torch.int32, device=dev) for _ in range(3))
    rows = api.IndexRows(start=out[0], end=out[1], order=out[2],
                         seg_out_end=seg_out_end[:len(owned)].cpu(),
                         seg_rec_id=[s.rec_id for s in owned], n=n, irregular=general,
                         replays=replays, general=general)
    n_local = torch.tensor([n], dtype=torch.int64, device=dev)
    counts, offset = ordered_offsets(n_local, group)
    return rows, counts, offset


def query_rows_for_range(f1, lo: int, hi: int, k_max: int):
    """Index rows a rank needs to answer queries with k <= k_max over window positions
    [lo, hi) of a record: only rows with lo < f1 <= hi + k_max - 1 can cover them (f2 >= f1;
    src/memo_query.py:25-27,46-49), i.e. its own rows plus a right halo of k_max - 1 positions
    (SURVEY 8e).  f1: the record's row starts in index order (torch tensor or numpy array).
    Returns the half-open row range (first, last); no collective is involved."""
    import numpy as _np
    a = f1.cpu().numpy() if isinstance(f1, torch.Tensor) else _np.asarray(f1)
    first = int(_np.searchsorted(a, lo, side="right"))
    last = int(_np.searchsorted(a, hi + max(k_max, 1) - 1, side="right"))
    return first, last
