#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL): every rank
builds the index rows of its position shard with shard.build_index_sharded, the
rows are gathered on rank 0 in rank order and compared with the oracle's
unsharded index -- for matching statistics and for irregular input (carry
exchange), conservation and membership."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from memo_b200 import api, shard  # noqa: E402
from oracle import memo_oracle as mo  # noqa: E402  (checker)


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for C in (9, 40, 93):
        lens = [30011, 1, 52000, 2500, 41234]
        recs = [(f"c{i}", n) for i, n in enumerate(lens)]
        total = sum(lens)
        valid = np.concatenate([mo.synth_dap(n, C, seed=900 + i, dense=(i % 2 == 0)) for i, n in enumerate(lens)])
        rng = np.random.default_rng(C)
        junk = rng.integers(0, 40, (total, C))
        junk[rng.random((total, C)) < 0.5] = 0
        junk[:, 0] = np.maximum(total + 5 - 2 * np.arange(total), 0)      # carries cross whole shards
        for name, vals in (("valid", valid), ("irregular", junk)):
            for order in (True, False):
                plan = shard.plan_shard(recs, total, world, rank, halo_right=64)
                buf = torch.from_numpy(np.ascontiguousarray(vals[plan.buf_lo:plan.buf_hi], dtype=np.int32)).to(dev)
                rows, counts, offset = shard.build_index_sharded(buf, plan, C, order)
                assert rows.general == (name == "irregular")
                cols = torch.stack([rows.start[:rows.n], rows.end[:rows.n], rows.order[:rows.n]])
                allrows = shard.gather_index_rows(cols, counts, dst=0)
                if rank == 0:
                    want = mo.index_build(vals, recs, order)
                    got = allrows.cpu().numpy()
                    same = (got.shape[1] == want[1].size and
                            np.array_equal(got[0].astype(np.int64), want[1]) and
                            np.array_equal(got[1].view(np.uint32).astype(np.int64), want[2]) and
                            np.array_equal(got[2].astype(np.int64), want[3]))
                    print(f"C={C} {name} order={order}: rows={got.shape[1]} offsets={counts.tolist()} "
                          f"{'OK' if same else 'MISMATCH'}", flush=True)
                    ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_check", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
