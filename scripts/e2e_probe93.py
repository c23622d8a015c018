"""Where does the end-to-end time of the 94-genome step go?  (run on the GPU box)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from memo_b200 import api, host
dev = torch.device("cuda", 0)
L, C, k = 16_000_000, 93, 31
REC = 248_956_422
dap = api.synth_dap(REC, C, 20240612, rows=L, device=dev)
h = torch.empty((L, C), dtype=torch.int32, pin_memory=True); h.copy_(dap); torch.cuda.synchronize()
d2 = torch.empty_like(dap)
for _ in range(2):
    t0 = time.perf_counter(); d2.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D pinned whole: %.1f ms  %.1f GB/s" % (dt * 1e3, h.numel() * 4 / dt / 1e9))
# chunked, back to back on one stream, destination shifted by one row (the halo row)
rows_c = (256 << 20) // (4 * C)
st = torch.cuda.Stream(dev)
for shift in (0, 1):
    bufs = [torch.empty((rows_c + 1, C), dtype=torch.int32, device=dev) for _ in range(3)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(st):
        for i, a in enumerate(range(0, L, rows_c)):
            n = min(rows_c, L - a)
            bufs[i % 3][shift:shift + n].copy_(h[a:a + n], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("H2D pinned in 256 MB chunks (dst shift %d row): %.1f ms  %.1f GB/s" % (shift, dt * 1e3, h.numel() * 4 / dt / 1e9))
# both directions at once
out = torch.empty((3, 30_000_000), dtype=torch.int32, device=dev)
pin_out = torch.empty((3, 30_000_000), dtype=torch.int32, pin_memory=True)
st2 = torch.cuda.Stream(dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(st):
    d2.copy_(h, non_blocking=True)
with torch.cuda.stream(st2):
    for _ in range(4):
        pin_out.copy_(out, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D 5.95 GB with 1.44 GB of D2H beside it: %.1f ms" % (dt * 1e3))
del d2, out, pin_out
segs = [api.Segment(row_begin=0, n_rows=L, pos0=0, rec_len=REC, rec_id=0, flags=1)]
for cb in (64 << 20, 256 << 20, 512 << 20):
    for _ in range(2):
        t0 = time.perf_counter()
        rows = host.build_index(h, None, True, device=dev, segs=segs, raw=True, chunk_bytes=cb)
        t1 = time.perf_counter()
        q = host.query(rows.start, rows.end, rows.order, 0, L, k, C + 1, False, device=dev, raw=True, trusted=True)
        t2 = time.perf_counter()
        del q
    print("chunk %d MB: build_index %.1f ms (%.1f GB/s), query %.1f ms, rows %d" % (
        cb >> 20, (t1 - t0) * 1e3, h.numel() * 4 / (t1 - t0) / 1e9, (t2 - t1) * 1e3, rows.n))
# finer: query pieces
t0 = time.perf_counter(); t1_, t2_, t3_ = host._rows_to_device(rows.start, rows.end, rows.order, dev, True); torch.cuda.synchronize(); a = time.perf_counter()
out = api.query_conservation(t1_, t2_, t3_, 0, L, k, C + 1); torch.cuda.synchronize(); b = time.perf_counter()
stg = host._pinned("qout", L)[:L]; stg.copy_(out, non_blocking=True); torch.cuda.synchronize(); c = time.perf_counter()
x = stg.numpy().copy(); d = time.perf_counter()
print("query: rows->dev %.1f ms, kernel %.1f ms, D2H %.1f ms, host copy %.1f ms" % ((a - t0) * 1e3, (b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3))
