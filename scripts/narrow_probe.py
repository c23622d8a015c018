"""Why is the 10-genome build slower after the 94-genome workload ran in the same process?  (GPU box)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from memo_b200 import _lib
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()

def narrow(tag):
    wl = bench.Workload(100_000_000, 9, 31, False, bench.SEED0 + 1, 0, 1, dev, {})
    t = bench.timed_run(wl, 10, 3, torch.cuda.synchronize, 0)
    print("%-46s kern %.3f ms idx %.3f ms  dap ptr %#x" % (tag, t["kern_ms"], t["idx_ms"], wl.dap.data_ptr()), flush=True)
    del wl
    torch.cuda.empty_cache()

narrow("fresh process")
x = torch.empty(92_000_000_000, dtype=torch.uint8, device=dev); x.fill_(1); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
narrow("after 92 GB allocated, written, freed")
import time
x = torch.empty(92_000_000_000, dtype=torch.uint8, device=dev); x.fill_(1); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
time.sleep(2.0)
narrow("after 92 GB allocated, written, freed + 2 s")
x = torch.empty(92_000_000_000, dtype=torch.uint8, device=dev); x.fill_(1); torch.cuda.synchronize(); del x; torch.cuda.empty_cache()
time.sleep(0.5)
narrow("after 92 GB allocated, written, freed + 0.5 s")
wl = bench.Workload(20_000_000, 93, 31, False, bench.SEED0, 0, 1, dev, {})
bench.timed_run(wl, 3, 1, torch.cuda.synchronize, 0); del wl; torch.cuda.empty_cache()
narrow("after a 94-genome x 20 Mbp workload")
os.environ["X"] = "1"
wl = bench.Workload(bench.CHR1, 93, 31, False, bench.SEED0, 0, 1, dev, {})
bench.timed_run(wl, 3, 1, torch.cuda.synchronize, 0); del wl; torch.cuda.empty_cache()
narrow("after the chr1 x 94 workload")
narrow("once more")
