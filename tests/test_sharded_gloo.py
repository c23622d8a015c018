"""The N > 1 path on CPU: world_size-2 (and 3) gloo process groups drive the
collectives of memo_b200/shard.py -- the ordered-offset all-gather, the carry
exchange of the exact build, the irregular verdict and the final ordered gather
of index rows -- with the C oracle standing in for the device builder (checker
only: the product function build_index_sharded refuses CPU tensors)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NONE32 = 0xFFFFFFFF


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _last_flagged_end(vals, pos0, primed, prev_row):
    """Per column: p + v at the last flagged row of a run (flag = first row of a
    primed run, or v[r-1] <= v[r]); NONE32 if nothing is flagged."""
    n, C = vals.shape
    out = np.full(C, NONE32, dtype=np.int64)
    prev = prev_row
    for r in range(n):
        for j in range(C):
            if (r == 0 and primed) or (prev is not None and prev[j] <= vals[r, j]):
                out[j] = pos0 + r + vals[r, j]
        prev = vals[r]
    return out


def _worker(rank, world, port, lens, C, seed, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from memo_b200 import shard
        from oracle import c_oracle as co
        from oracle import memo_oracle as mo

        recs = [(f"c{i}", l) for i, l in enumerate(lens)]
        total = sum(lens)
        vals = np.concatenate([mo.synth_dap(n, C, seed=seed + i, dense=True) for i, n in enumerate(lens)])
        plan = shard.plan_shard(recs, total, world, rank, halo_right=40)

        # ---- valid matching statistics: shards are independent; ordered offsets + gather
        segs = [co.Seg(s.row_begin, s.n_rows, s.pos0, s.rec_len, s.rec_id, s.flags)
                for s in plan.segs[:plan.n_owned]]
        part = co.index_build(vals[plan.buf_lo:plan.buf_hi], recs, True, segs=segs)
        cols = torch.from_numpy(np.stack([part[1], part[2], part[3]]).astype(np.int32))
        counts, offset = shard.ordered_offsets(torch.tensor([cols.shape[1]], dtype=torch.int64))
        assert int(counts[rank]) == cols.shape[1]
        assert int(offset) == int(counts[:rank].sum())
        allrows = shard.gather_index_rows(cols, counts, dst=0)
        if rank == 0:
            whole = co.index_build(vals, recs, True)
            got = allrows.numpy().astype(np.int64)
            for j in range(3):
                assert np.array_equal(got[j], whole[j + 1]), "gathered rows differ from the unsharded index"
        else:
            assert allrows is None

        # ---- irregular input: verdict + carry exchange
        rng = np.random.default_rng(seed)
        junk = rng.integers(0, 6, size=(total, C)).astype(np.int64)      # not matching statistics
        junk[rng.random((total, C)) < 0.7] = 0
        # make long stretches without any flagged row so that carries cross whole shards
        junk[:, 0] = np.maximum(total + 5 - np.arange(total) * 2, 0)
        assert shard.agree_irregular(rank == world - 1, torch.device("cpu")) is True
        assert shard.agree_irregular(False, torch.device("cpu")) is False
        owned = plan.segs[:plan.n_owned]
        carry = np.full(C, NONE32, dtype=np.int64)
        for s in owned:                                                    # what the kernel hands on
            lo = plan.buf_lo + s.row_begin
            primed = bool(s.flags & 1)
            prev = None if primed else junk[lo - 1]
            c = _last_flagged_end(junk[lo:lo + s.n_rows], s.pos0, primed, prev)
            carry = np.where(c != NONE32, c, carry) if not primed else c
        cin = shard.exchange_carries(torch.from_numpy(carry)).numpy()
        if owned and not (owned[0].flags & 1):
            # brute force: last flagged end among the record's rows before this shard
            s0 = owned[0]
            rec_first = plan.pos_lo - s0.pos0
            want = _last_flagged_end(junk[rec_first:plan.pos_lo], 0, True, None)
            assert np.array_equal(cin, want), (rank, cin, want)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:                                    # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("world,lens", [(2, [3000, 1, 4500, 2500]), (3, [9000]), (2, [5, 7])])
def test_sharded_collectives_gloo(world, lens):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, 5, 77, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_carry_in_from_gathered():
    from memo_b200 import shard
    c = torch.tensor([[5, NONE32, 7], [NONE32, NONE32, 9], [1, NONE32, NONE32]], dtype=torch.int64)
    assert shard.carry_in_from_gathered(c, 0).tolist() == [NONE32] * 3
    assert shard.carry_in_from_gathered(c, 1).tolist() == [5, NONE32, 7]
    assert shard.carry_in_from_gathered(c, 2).tolist() == [5, NONE32, 9]


def test_build_index_sharded_refuses_cpu():
    from memo_b200 import shard, _lib
    plan = shard.plan_shard([("a", 10)], 10, 1, 0)
    with pytest.raises(_lib.MemoError):
        shard.build_index_sharded(torch.zeros((10, 3), dtype=torch.int32), plan, 3, True)


def _text_worker(rank, world, port, dap_path, out_path, lens, order, q):
    """A rank of the sharded dap_to_bed, with the C oracle standing in for the device build:
    byte share of dap.txt + the line before it -> BED part -> ordered_file_write."""
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from memo_b200 import io, shard
        from oracle import c_oracle as co
        from oracle import memo_oracle as mo
        recs = [(f"c{i}", l) for i, l in enumerate(lens)]
        lo, hi, prev = io.split_text_rows(dap_path, world, rank)
        blocks = list(io.iter_dap_text(dap_path, block_bytes=1 << 14, byte_range=(lo, hi)))
        part = f"{out_path}.part{rank:03d}"
        with open(part, "w") as fh:
            if blocks:
                pos0 = blocks[0][0]
                vals = np.concatenate([b for _, b in blocks])
                starts = set(np.cumsum([0] + lens[:-1]).tolist())
                halo = prev is not None and pos0 not in starts
                buf = np.vstack([[int(x) for x in prev.split()][1:], vals]) if halo else vals
                segs = co.make_segs(recs, len(vals), pos_first=pos0, row0=1 if halo else 0,
                                    primed_first=not halo, chr_end_last=True)
                tail = segs[-1]
                if rank != world - 1 and tail.pos0 + tail.n_rows < tail.rec_len:
                    tail.flags &= ~2
                fh.write(mo.format_bed(recs, *co.index_build(buf, recs, order, segs=segs)))
        shard.ordered_file_write(part, out_path, torch.device("cpu"))
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:                                         # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_text_shares_and_ordered_write_gloo(world, tmp_path):
    from oracle import c_oracle as co
    from oracle import memo_oracle as mo
    lens = [4000, 1, 2500, 3300]
    recs = [(f"c{i}", l) for i, l in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, 4, seed=31 + i, dense=True) for i, n in enumerate(lens)])[:-55]
    dap = tmp_path / "dap.txt"
    dap.write_text("".join(f"{i} " + " ".join(map(str, row)) + "\n" for i, row in enumerate(vals)))
    out = tmp_path / "out.bed"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_text_worker, args=(r, world, port, str(dap), str(out), lens, True, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
    assert out.read_text() == mo.format_bed(recs, *co.index_build(vals, recs, True))
    assert not list(tmp_path.glob("out.bed.part*"))
