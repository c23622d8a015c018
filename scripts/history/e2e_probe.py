"""Where does the end-to-end time go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from memo_b200 import api, host
dev = torch.device("cuda", 0)
L, C, k = 100_000_000, 9, 31
dap = api.synth_dap(L, C, 20240612, device=dev)
h = torch.empty((L, C), dtype=torch.int32, pin_memory=True); h.copy_(dap); torch.cuda.synchronize()
d2 = torch.empty_like(dap)
for _ in range(2):
    t0 = time.perf_counter(); d2.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D pinned whole: %.1f ms  %.1f GB/s" % (dt * 1e3, h.numel() * 4 / dt / 1e9))
del d2
segs = [api.Segment(row_begin=0, n_rows=L, pos0=0, rec_len=L, rec_id=0, flags=3)]
for cb in (16 << 20, 64 << 20, 256 << 20):
    for _ in range(2):
        t0 = time.perf_counter()
        rows = host.build_index(h, None, True, device=dev, segs=segs, raw=True, chunk_bytes=cb)
        t1 = time.perf_counter()
        q = host.query(rows.start, rows.end, rows.order, 0, L, k, C + 1, False, device=dev, raw=True, trusted=True)
        t2 = time.perf_counter()
    print("chunk %d MB: build_index %.1f ms, query %.1f ms, rows %d" % (cb >> 20, (t1 - t0) * 1e3, (t2 - t1) * 1e3, rows.n))
# finer: query pieces
t0 = time.perf_counter(); t1_, t2_, t3_ = host._rows_to_device(rows.start, rows.end, rows.order, dev, True); torch.cuda.synchronize(); a = time.perf_counter()
out = api.query_conservation(t1_, t2_, t3_, 0, L, k, C + 1); torch.cuda.synchronize(); b = time.perf_counter()
st = host._pinned("qout", L)[:L]; st.copy_(out, non_blocking=True); torch.cuda.synchronize(); c = time.perf_counter()
x = st.numpy().copy(); d = time.perf_counter()
print("query: rows->dev %.1f ms, kernel %.1f ms, D2H %.1f ms, host copy %.1f ms" % ((a - t0) * 1e3, (b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3))
