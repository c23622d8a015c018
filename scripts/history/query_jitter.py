"""Distribution of the query kernel time over many launches.  usage: query_jitter.py cols rows [membership]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from memo_b200 import api, _lib
C, L = int(sys.argv[1]), int(sys.argv[2])
memb = len(sys.argv) > 3
dap = api.synth_dap(L, C, seed=20240612)
res = api.index_build(dap, [("chrS", L)], not memb)
n = res.n
ws = torch.zeros(_lib.load().memo_query_workspace_bytes(L), dtype=torch.uint8, device="cuda")
st = torch.zeros(1, dtype=torch.int32, device="cuda")
out = None
ts = []
for it in range(60):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if memb:
        out = api.query_membership(res.start[:n], res.end[:n], res.order[:n], 0, L, 31, C + 1, out=out, check=False, status=st, workspace=ws)
    else:
        out = api.query_conservation(res.start[:n], res.end[:n], res.order[:n], 0, L, 31, C + 1, out=out, check=False, status=st, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts = ts[5:]
print(os.environ.get("MEMO_QUERY_HEAVY", "default"), "us: min %.0f median %.0f max %.0f" % (min(ts), sorted(ts)[len(ts) // 2], max(ts)), " >1.5x median:", sum(t > 1.5 * sorted(ts)[len(ts) // 2] for t in ts), [int(t) for t in ts[:20]])
