"""Position-range sharding of the pivot across the GPUs of one box.

Shards are independent: rank r builds the index rows of its own position range
from its own DAP rows plus a one-row left halo (the flag compares a row with its
predecessor) and, for queries, a right halo of k_max-1 rows.  The only exchange
is an all-gather of one int64 per rank -- the number of index rows a rank owns --
whose exclusive prefix is the rank's offset in the ordered index.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from ._lib import MEMO_SEG_CHR_END, MEMO_SEG_PRIMED, Segment


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
    # a run that continues on the next rank leaves its chr-end rows to that rank
    if owned and hi < total_rows:
        last = owned[-1]
        rec_end = sum(l for _, l in records[:last.rec_id + 1])
        if hi < rec_end:
            last.flags &= ~MEMO_SEG_CHR_END
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
