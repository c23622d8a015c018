"""Small repro of the strip build at 93 columns against the oracle (debug aid)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memo_b200 import api
from oracle import memo_oracle as mo

L, C = int(os.environ.get("L", 3000)), int(os.environ.get("C", 93))
dap = api.synth_dap(L, C, seed=20240614)
host = dap.cpu().numpy()
recs = [("chrA", L)]
for order in (True, False):
    res = api.index_build(dap, recs, order, kernel_variant=3)
    got = res.to_host()
    want = mo.index_build(host, recs, order)
    print("order", order, "n", res.n, "want", want[1].size, "equal", all(np.array_equal(g, w) for g, w in zip(got, want)))
