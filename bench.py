#!/usr/bin/env python3
"""bench.py -- MEMO hot path on B200: pivot bp/s for conservation index build +
k=31 window query on BASELINE.json configs[3] -- a chr1-sized pivot record
(248 956 422 bp) x 94 genomes, synthetic HPRC-shaped DAP -- on 1 GPU, or cut into
N position shards on N GPUs (strong scaling), with the HBM roofline of the
dominant kernel, an end-to-end leg through the host-buffer API, parity against
the C port of the reference algorithm, and that port timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm (oracle port, all host cores)

One "step" = one pass of the hot path over the workload: index build of the
rank's DAP shard (device resident) followed by the k-mer conservation query over
the shard's window on the freshly built index rows.  value = total pivot bp over
all ranks / max-over-ranks device time.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED0 = 20240611               # seed = 20240611 + config index
KH = 128                       # right halo rows kept for queries (k <= 129)
CHR1 = 248_956_422


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(key):
    """DRAM bytes (read + write) per launch from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json, scripts/ncu_traffic.py), or
    None.  A profile lookup, not a measurement of this run: the source is named
    in the line."""
    try:
        data = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return data.get(key), data.get("_source")
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle's C port on host cores (the reference itself is Python and
# does not exist on the GPU box; see DESIGN.md "measurement")
# ---------------------------------------------------------------------------
def cpu_port_run(dap_np, rec_len, row0, k, n_docs, threads, order=True, halo_first=False, do_query=True):
    """Index build + k-mer query over rows [row0, row0 + n) of one record held in
    `dap_np` (with halo_first, dap_np[0] is row row0 - 1, the halo of a slice that
    starts inside the record), cut into `threads` slices with a one-row halo
    (exact for matching statistics).  Returns (seconds, n_out, query, rows)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle as co

    off = 1 if halo_first else 0
    n, C = dap_np.shape[0] - off, dap_np.shape[1]
    recs = [("chrS", rec_len)]
    cuts = [row0 + (n * i) // threads for i in range(threads + 1)]

    def work(i):
        a, b = cuts[i], cuts[i + 1]
        if b <= a:
            return tuple(np.zeros(0, dtype=np.int64) for _ in range(4))
        lo = a - 1 if a > 0 else a          # halo row (the record's previous row)
        segs = co.make_segs(recs, b - a, pos_first=a, row0=a - lo, primed_first=(a == 0),
                            chr_end_last=(b == rec_len))
        cap = int((b - a) * C * 0.06) + 4 * C
        while True:
            r = co.index_build(dap_np[lo - row0 + off:b - row0 + off], recs, order, segs=segs, cap=cap)
            if r[1].size < cap:
                return r
            cap *= 2

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, range(threads)))
    f1 = np.concatenate([p[1] for p in parts])
    f2 = np.concatenate([p[2] for p in parts])
    f3 = np.concatenate([p[3] for p in parts])

    def qwork(i):
        return co.query(f1, f2, f3, cuts[i], cuts[i + 1], k, n_docs, not order)

    outs = None
    if do_query:
        with ThreadPoolExecutor(threads) as ex:
            outs = np.concatenate(list(ex.map(qwork, range(threads))))
    dt = time.perf_counter() - t0
    return dt, int(f1.size), outs, (f1, f2, f3)


def workload_config(args, n_gpus):
    memb = args.membership
    rows, genomes = args.rows, args.cols + 1
    if args.cols == 93 and rows == CHR1 and not memb:
        tag = "BASELINE configs[3]: HPRC-shaped chr1 (248 956 422 bp pivot) x 94 genomes"
    elif args.cols == 9 and rows == 100_000_000 and not memb:
        tag = "BASELINE configs[1]: 10 genomes x 100 Mbp pivot"
    elif memb:
        tag = "BASELINE configs[2]-shaped: membership index (-m) and per-genome presence bitmaps"
    else:
        tag = "non-default shape"
    what = "membership index build + membership" if memb else "conservation index build +"
    return {"workload": f"{tag}; synthetic HPRC-shaped DAP, {genomes} genomes x {rows} bp pivot record "
                        f"cut into {n_gpus} position shard(s), {what} k={args.k} window query over "
                        "every shard",
            "genomes": genomes, "pivot_bp": rows, "pivot_bp_per_gpu": rows // n_gpus, "k": args.k,
            "partition": f"position ranges x{n_gpus} of one record, 1-row left halo, {KH}-row right halo",
            "l2": f"inputs ({rows // n_gpus * args.cols * 4 / 1e9:.2f} GB DAP per GPU) are larger than L2; "
                  "no flush needed",
            "seed": seed_of(args)}


def seed_of(args):
    if args.cols == 9 and not args.membership:
        return SEED0 + 1
    if args.membership:
        return SEED0 + 2
    return SEED0 + 3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as co
    C, k = args.cols, args.k
    cores = os.cpu_count() or 1
    sample = min(args.rows, args.ref_sample_rows)
    dap = co.synth_dap(args.rows, C, seed_of(args), row0=0, rows=sample, threads=cores)
    times = []
    for i in range(args.warmup + args.steps):
        dt, n_out, _, _ = cpu_port_run(dap, args.rows, 0, k, C + 1, cores, order=not args.membership)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": "pivot bp/s (%s index build + k-mer query)" %
                                       ("membership" if args.membership else "conservation"),
        "value": value, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "bp/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} rows of the workload per step; C port of the "
                                   "reference algorithm (oracle/memo_oracle.c), one slice per core; the "
                                   "reference itself is Python (no compiled source to build) and is "
                                   "absent on this box: BASELINE.md records its rate"},
        "e2e": {"value": value, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
class Workload:
    """One record of `rec_len` rows x C columns, position-sharded over the ranks;
    this rank's shard resident on its device, outputs sized by a counting run."""

    def __init__(self, rec_len, C, k, membership, seed, rank, world, dev, tuning, span=None):
        import torch
        from memo_b200 import api, shard, _lib
        self.torch, self.api = torch, api
        self.rec_len, self.C, self.k, self.membership = rec_len, C, k, membership
        self.n_docs = C + 1
        self.rank, self.world, self.dev, self.tuning, self.seed = rank, world, dev, tuning, seed
        # the rank's position shard of the record, or an explicit span of it (whole-genome mode)
        self.lo, self.hi = shard.shard_range(rec_len, world, rank) if span is None else span
        self.Lr = self.hi - self.lo
        self.buf_lo = self.lo - 1 if self.lo > 0 else self.lo      # 1-row left halo
        self.buf_hi = min(self.hi + KH, rec_len)                   # right halo for the query
        self.dap = api.synth_dap(rec_len, C, seed, row0=self.buf_lo, rows=self.buf_hi - self.buf_lo, device=dev)
        segs = [api.Segment(row_begin=self.lo - self.buf_lo, n_rows=self.Lr, pos0=self.lo, rec_len=rec_len,
                            rec_id=0, flags=(api.MEMO_SEG_PRIMED if self.lo == 0 else 0) |
                                            (api.MEMO_SEG_CHR_END if self.hi == rec_len else 0))]
        if self.buf_hi > self.hi:
            segs.append(api.Segment(row_begin=self.hi - self.buf_lo, n_rows=self.buf_hi - self.hi,
                                    pos0=self.hi, rec_len=rec_len, rec_id=0, flags=0))
        self.segs = segs
        self.order = not membership
        self.builder = api.IndexBuilder(dev)
        self.seg_out_end = torch.zeros(len(segs), dtype=torch.int64, device=dev)
        self.builder.launch(self.dap, C, segs, self.order, None, self.seg_out_end, **tuning)   # counting run
        self.n_all, irregular, _ = self.builder.result()
        assert not irregular, "synthetic DAP must be valid matching statistics"
        self.out = tuple(torch.empty(self.n_all + 16, dtype=torch.int32, device=dev) for _ in range(3))
        self.q_out = torch.empty((self.Lr, (self.n_docs + 31) // 32), dtype=torch.int32, device=dev) \
            if membership else torch.empty(self.Lr, dtype=torch.uint8, device=dev)
        self.counts = torch.zeros(world, dtype=torch.int64, device=dev)
        self.q_status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.q_ws = torch.empty(max(_lib.load().memo_query_workspace_bytes(self.Lr), 1), dtype=torch.uint8,
                                device=dev)

    def build(self):
        self.builder.launch(self.dap, self.C, self.segs, self.order, self.out, self.seg_out_end, **self.tuning)

    def query(self):
        n, o, api = self.n_all, self.out, self.api
        fn = api.query_membership if self.membership else api.query_conservation
        fn(o[0][:n], o[1][:n], o[2][:n], self.lo, self.hi, self.k, self.n_docs, out=self.q_out,
           check=False, status=self.q_status, workspace=self.q_ws)

    def step(self, ev=None):
        # nothing in here waits for the device: the row count is known from the sizing
        # run (and re-checked after the timed region), the query reads the fresh rows
        import torch.distributed as dist
        if ev:
            ev[0].record()
        self.build()
        if ev:
            ev[1].record()
        if self.world > 1:                               # ordered write offsets: gather the counts
            dist.all_gather_into_tensor(self.counts, self.seg_out_end[:1])
        if ev:
            ev[2].record()
        self.query()
        if ev:
            ev[3].record()

    def check_invariant(self):
        """conservation == 1 + #{MS >= k} / membership == [1, MS >= k] over the whole shard
        (holds for valid matching statistics; independent of the build)."""
        torch, np = self.torch, __import__("numpy")
        dap, off = self.dap, self.lo - self.buf_lo
        if self.membership:
            n_chk = min(self.Lr, 2_000_000)
            bits = self.api.unpack_membership(self.q_out[:n_chk].cpu().numpy(), self.n_docs)
            want = (dap[off:off + n_chk] >= self.k).cpu().numpy().astype(np.uint8)
            assert bits[:, 0].all() and np.array_equal(bits[:, 1:], want), "membership != [1, MS >= k]"
            return
        step_rows = 4_000_000                            # (in slices: the comparison needs temporaries)
        for a in range(0, self.Lr, step_rows):
            b = min(a + step_rows, self.Lr)
            want = (1 + (dap[off + a:off + b] >= self.k).sum(dim=1)).to(torch.uint8)
            assert torch.equal(self.q_out[a:b], want), "query result violates conservation == 1 + #{MS >= k}"
            del want

    def owned_rows_between(self, p_lo, p_hi):
        """Index rows (start, end, order as int64 numpy) this rank owns with p_lo <= start < p_hi."""
        import numpy as np
        torch = self.torch
        n = int(self.seg_out_end[0].item())
        f1 = self.out[0][:n]
        a = int(torch.searchsorted(f1, torch.tensor([p_lo], dtype=torch.int32, device=self.dev)).item())
        b = int(torch.searchsorted(f1, torch.tensor([p_hi], dtype=torch.int32, device=self.dev)).item())
        return (self.out[0][a:b].cpu().numpy().astype(np.int64),
                self.out[1][a:b].cpu().numpy().view(np.uint32).astype(np.int64),
                self.out[2][a:b].cpu().numpy().astype(np.int64))


def timed_run(wl, steps, warmup, barrier, local):
    """W warm-up steps, then exactly `steps` steps bracketed by barrier + synchronize."""
    import torch
    from memo_b200 import _lib
    lib = _lib.load()
    for _ in range(warmup):
        wl.step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    lib.memo_profile_enable(1)          # CUDA events around the streaming kernel of every build
    lib.memo_launch_count(1)
    t_begin.record()
    for i in range(steps):
        wl.step(evs[i])
    t_end.record()
    barrier()
    launches = int(lib.memo_launch_count(1))
    n_out, irregular, _ = wl.builder.result()
    assert n_out == wl.n_all and not irregular and int(wl.q_status.item()) == 0
    parked = int(wl.builder._result[3].item())           # (stat of the single-kernel strip build)
    lib.memo_profile_enable(0)
    k_ms, k_n = ctypes.c_double(0.0), ctypes.c_int32(0)
    _lib.check(lib.memo_profile_collect(ctypes.byref(k_ms), ctypes.byref(k_n)), "memo_profile_collect")
    clocks = sampler.stop()
    return {"total_ms": t_begin.elapsed_time(t_end),
            "idx_ms": sum(e[0].elapsed_time(e[1]) for e in evs) / steps,
            "qry_ms": sum(e[2].elapsed_time(e[3]) for e in evs) / steps,
            "kern_ms": k_ms.value / max(k_n.value, 1), "launches": launches, "clocks": clocks,
            "parked_rows": parked}


def rooflines(wl, t, peak, peak_src, n_rows_shard, tag):
    """Roofline objects of one rank's shard (SURVEY 8d algorithmic bytes)."""
    C, Lr = wl.C, wl.Lr
    bytes_idx = 4.0 * Lr * C + 12.0 * n_rows_shard       # DAP read once + index rows written once
    ach = bytes_idx / (t["kern_ms"] * 1e-3) / 1e9
    ach_build = bytes_idx / (t["idx_ms"] * 1e-3) / 1e9
    bytes_q = 12.0 * wl.n_all + (4.0 * ((wl.n_docs + 31) // 32) if wl.membership else 1.0) * Lr
    ach_q = bytes_q / (t["qry_ms"] * 1e-3) / 1e9
    kname = "wide2_kernel" if wl.tuning.get("kernel_variant") == 3 else ("narrow_kernel" if C <= 16 else "wide_kernel")
    tr_k, src = profiled_traffic(f"{tag}_stream_kernel")
    tr_b, _ = profiled_traffic(f"{tag}_index_build")
    tr_q, _ = profiled_traffic(f"{tag}_query")
    return {
        "roofline": {"kernel": kname + " (streaming kernel of memo_index_build: DAP -> index rows; CUDA events "
                                       "around the launch on its stream, averaged over the timed steps)",
                     "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": tr_k, "traffic_source": src, "peak_source": peak_src,
                     "algorithmic_bytes": bytes_idx, "kernel_ms": t["kern_ms"]},
        "roofline_index_build": {"kernel": "memo_index_build (whole call)", "bound": "hbm",
                                 "achieved": ach_build, "peak": peak, "unit": "GB/s", "frac": ach_build / peak,
                                 "algorithmic_bytes": bytes_idx, "traffic": tr_b},
        "roofline_query": {"kernel": "query_planes_kernel (" + ("membership" if wl.membership else "conservation") + ")",
                           "bound": "hbm", "achieved": ach_q, "peak": peak, "unit": "GB/s", "frac": ach_q / peak,
                           "algorithmic_bytes": bytes_q, "traffic": tr_q},
    }


def shard_parity(wl, args):
    """Driver-visible multi-GPU parity: for every cut between two position shards,
    the right-hand rank collects the index rows either side of the cut (its own first
    `m` positions and, point to point, the left-hand rank's last `m`), and compares
    their concatenation in rank order with the C port of the reference algorithm run
    over the 2m positions spanning the cut; the rows' places in the ordered index
    (exclusive prefix of the all-gathered counts) are checked on the way."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from memo_b200 import shard
    world, rank, dev = wl.world, wl.rank, wl.dev
    m = min(args.parity_rows, wl.rec_len // world)       # the same on every rank
    n_local = wl.seg_out_end[:1].clone()
    counts, offset = shard.ordered_offsets(n_local)
    counts_h = counts.cpu().tolist()
    ok = int(offset.item()) == sum(counts_h[:rank]) and counts_h[rank] == int(n_local.item())
    # left-hand side of the cut at wl.hi goes to rank + 1
    if rank + 1 < world:
        s, e, o = wl.owned_rows_between(wl.hi - m, wl.hi)
        t = torch.from_numpy(np.stack([s, e, o])).to(dev)
        dist.send(torch.tensor([t.shape[1]], dtype=torch.int64, device=dev), dst=rank + 1)
        dist.send(t.contiguous(), dst=rank + 1)
    if rank > 0:
        n_left = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.recv(n_left, src=rank - 1)
        left = torch.empty((3, int(n_left.item())), dtype=torch.int64, device=dev)
        dist.recv(left, src=rank - 1)
        left = left.cpu().numpy()
        cut = wl.lo
        right = wl.owned_rows_between(cut, cut + m)
        # the 2m positions spanning the cut, generated afresh (plus the halo row before them)
        p0 = cut - m                              # >= 0: every shard holds at least m positions
        h = 1 if p0 > 0 else 0
        rows = wl.api.synth_dap(wl.rec_len, wl.C, wl.seed, row0=p0 - h, rows=2 * m + h, device=dev)
        _, _, _, want = cpu_port_run(rows.cpu().numpy(), wl.rec_len, p0, wl.k, wl.n_docs,
                                     min(os.cpu_count() or 1, 8), order=wl.order, halo_first=bool(h),
                                     do_query=False)
        sel = want[0] < cut + m                   # (chr-end rows of a shard that ends the record start at rec_len)
        got = [np.concatenate([left[i], right[i]]) for i in range(3)]
        ok = ok and all(np.array_equal(g, w[sel]) for g, w in zip(got, want))
        ok = ok and int((left[0] >= cut).sum()) == 0 and (right[0].size == 0 or int(right[0].min()) >= cut)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item()), m


def e2e_leg(wl, args, barrier):
    """The same step through the host-buffer API (host.build_index + host.query) on a
    host-resident slice of the shard: H2D of the DAP, D2H of rows + result inside the
    timed wall-clock region."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from memo_b200 import api, host
    n = min(wl.Lr, args.e2e_rows)
    off = wl.lo - wl.buf_lo
    sl_lo, sl_hi = wl.lo, wl.lo + n
    halo = min(KH, wl.buf_hi - sl_hi)
    view = wl.dap[0:off + n + halo]
    host_dap = torch.empty(tuple(view.shape), dtype=torch.int32, pin_memory=True)
    host_dap.copy_(view)
    ends_record = sl_hi == wl.rec_len
    segs = [api.Segment(row_begin=off, n_rows=n, pos0=sl_lo, rec_len=wl.rec_len, rec_id=0,
                        flags=(api.MEMO_SEG_PRIMED if sl_lo == 0 else 0) |
                              (api.MEMO_SEG_CHR_END if ends_record else 0))]
    if halo:
        segs.append(api.Segment(row_begin=off + n, n_rows=halo, pos0=sl_hi, rec_len=wl.rec_len, rec_id=0, flags=0))
    h2d = d2h = 0

    def one():
        rows_ = host.build_index(host_dap, None, wl.order, device=wl.dev, segs=segs, raw=True, **wl.tuning)
        q_ = host.query(rows_.start, rows_.end, rows_.order, sl_lo, sl_hi, wl.k, wl.n_docs, wl.membership,
                        device=wl.dev, raw=True, trusted=True)
        return rows_, q_

    rows_, q_ = one()          # one untimed pass: pinned result blocks and device buffers get allocated here
    del rows_, q_
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        rows_ = q_ = None      # drop the previous step's results (their pinned blocks recycle)
        rows_, q_ = one()
        h2d += host_dap.numel() * 4 + 12 * rows_.n
        d2h += 12 * rows_.n + q_.size
    torch.cuda.synchronize()
    if wl.world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=wl.dev)
    tot = torch.tensor([float(n)], dtype=torch.float64, device=wl.dev)
    if wl.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    # what the host can feed at this rank count, with nothing else going on: plain pinned -> device
    # copies of the same buffer on every rank at once (the ceiling the e2e figure runs against)
    probe = torch.empty(host_dap.numel() // 2, dtype=torch.int32, device=wl.dev)
    flat = host_dap.view(-1)[:probe.numel()]
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        probe.copy_(flat, non_blocking=True)
    torch.cuda.synchronize()
    if wl.world > 1:
        dist.barrier()
    h2d_ceiling = 4 * probe.numel() * 4 / (time.perf_counter() - t0) / 1e9
    del probe
    # same rows as the device-resident path (the slice's rows are a prefix of the shard's)
    s, e, o = wl.owned_rows_between(sl_lo, sl_hi if not ends_record else wl.rec_len + 1)
    if not wl.membership:
        ok = (np.array_equal(rows_.start[:s.size].astype(np.int64), s) and
              np.array_equal(rows_.end[:s.size].astype(np.int64), e) and
              np.array_equal(rows_.order[:s.size].astype(np.int64), o) and
              np.array_equal(q_, wl.q_out[:n].cpu().numpy()))
        assert ok, "host-buffer path disagrees with the device-resident path"
    return {"value": tot.item() * args.e2e_steps / t.item(), "unit": "bp/s",
            "h2d_bytes_per_step": h2d // args.e2e_steps, "d2h_bytes_per_step": d2h // args.e2e_steps,
            "steps": args.e2e_steps, "ms_per_step": 1e3 * t.item() / args.e2e_steps,
            "pivot_bp_per_gpu": n,
            "h2d_gbs_per_gpu": host_dap.numel() * 4 * args.e2e_steps / t.item() / 1e9,
            "h2d_ceiling_gbs_per_gpu": h2d_ceiling,
            "note": f"host.build_index + host.query on the first {n} positions of every shard: pinned host DAP "
                    "streamed to the device in chunks overlapped with the build, index rows (12 B each) and "
                    "the query result copied back to host; wall clock, max over ranks.  h2d_ceiling = plain "
                    "pinned -> device copies on all ranks at once (what the host can feed: the e2e figure "
                    "is bound by it, not by a collective)"}


def e2e_text_leg(wl, args):
    """The drop-in proper, from text: a dap.txt slice of the workload through
    memo_b200.dap_to_bed (file bytes -> pinned blocks -> device text parse -> device build ->
    BED text), wall clock."""
    import tempfile
    import numpy as np
    import pyarrow as pa
    import pyarrow.csv as pacsv
    from memo_b200 import dap_to_bed
    n = min(wl.Lr, args.e2e_text_rows)
    vals = wl.dap[:n].cpu().numpy()
    with tempfile.TemporaryDirectory() as d:
        dap, fai, bed = (os.path.join(d, f) for f in ("dap.txt", "p.fa.fai", "out.bed"))
        cols = [pa.array(np.arange(n, dtype=np.int64))] + [pa.array(vals[:, j]) for j in range(wl.C)]
        pacsv.write_csv(pa.table(cols, names=[f"f{i}" for i in range(wl.C + 1)]), dap,
                        write_options=pacsv.WriteOptions(include_header=False, delimiter=" "))
        with open(fai, "w") as fh:
            fh.write(f"chrS\t{wl.rec_len}\t6\t{wl.rec_len}\t{wl.rec_len + 1}\n")
        argv = ["--mem", "--overlap", "--fai", fai, "--dap", dap, "--out", bed] + (["--order"] if wl.order else [])
        a = dap_to_bed.parse_arguments(argv)
        dap_to_bed.check_args(a)
        dap_to_bed.main(a)                                 # (first call: pinned buffers get allocated)
        t0 = time.perf_counter()
        dap_to_bed.main(a)
        dt = time.perf_counter() - t0
        size, out_size = os.path.getsize(dap), os.path.getsize(bed)
    return {"value": n / dt, "unit": "bp/s", "rows": n, "seconds": dt, "dap_txt_bytes": size, "bed_bytes": out_size,
            "text_gbs": size / dt / 1e9,
            "note": "python -m memo_b200.dap_to_bed on a dap.txt slice of the workload, in process, wall clock: "
                    "file bytes -> pinned blocks -> device text parse (memo_dap_text_parse) -> streaming device "
                    "build -> BED text (pyarrow); the reference's dap_to_bed.py runs this at 32 kbp/s on one core "
                    "(BASELINE.md 2b)"}


def e2e_lengths_leg(wl, vals):
    """the same slice as e2e_text, but from per-genome MONI `.lengths.vert`
    files (index.sh:79) through `dap_to_bed --lengths` -- no `paste | nl`, no dap.txt.  The files
    are tokenized by host threads inside the library (memo_lengths_block_parse), the blocks go
    through the streaming device build.  The BED must equal the one the dap.txt route writes."""
    import filecmp
    import tempfile
    import numpy as np
    import pyarrow as pa
    import pyarrow.csv as pacsv
    from memo_b200 import dap_to_bed
    n = vals.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fai, bed, bed_text, dap = (os.path.join(d, f) for f in ("p.fa.fai", "out.bed", "text.bed", "dap.txt"))
        paths = []
        for j in range(wl.C):
            paths.append(os.path.join(d, f"g{j:03d}.w_rc.lengths.vert"))
            pacsv.write_csv(pa.table([pa.array(vals[:, j])], names=["v"]), paths[-1],
                            write_options=pacsv.WriteOptions(include_header=False))
        with open(fai, "w") as fh:
            fh.write(f"chrS\t{wl.rec_len}\t6\t{wl.rec_len}\t{wl.rec_len + 1}\n")
        flags = ["--mem", "--overlap", "--fai", fai] + (["--order"] if wl.order else [])
        a = dap_to_bed.parse_arguments(flags + ["--out", bed, "--lengths"] + paths)
        dap_to_bed.check_args(a)
        dap_to_bed.main(a)
        t0 = time.perf_counter()
        dap_to_bed.main(a)
        dt = time.perf_counter() - t0
        size = sum(os.path.getsize(p) for p in paths)
        # the same rows through the dap.txt route, for the comparison of the two BED files
        m = min(n, 100_000)
        cols = [pa.array(np.arange(m, dtype=np.int64))] + [pa.array(vals[:m, j]) for j in range(wl.C)]
        pacsv.write_csv(pa.table(cols, names=[f"f{i}" for i in range(wl.C + 1)]), dap,
                        write_options=pacsv.WriteOptions(include_header=False, delimiter=" "))
        small = []
        for j in range(wl.C):
            small.append(os.path.join(d, f"s{j:03d}.lengths"))
            with open(small[-1], "w") as fh:                 # MONI's own layout: header + one line
                fh.write(">chrS\n" + " ".join(map(str, vals[:m, j].tolist())) + "\n")
        for argv, out in ((["--dap", dap], bed_text), (["--lengths"] + small, bed)):
            b = dap_to_bed.parse_arguments(flags + ["--out", out] + argv)
            dap_to_bed.check_args(b)
            dap_to_bed.main(b)
        same = filecmp.cmp(bed, bed_text, shallow=False)
    return {"value": n / dt, "unit": "bp/s", "rows": n, "seconds": dt, "lengths_bytes": size, "files": wl.C,
            "text_gbs": size / dt / 1e9, "host_threads": min(16, os.cpu_count() or 1, wl.C),
            "equals_dap_txt_route": same, "rows_compared": m,
            "note": "python -m memo_b200.dap_to_bed --lengths on per-genome .lengths.vert files of the same slice, "
                    "in process, wall clock: files -> host tokenizer threads (memo_lengths_block_parse) -> "
                    "streaming device build -> BED text; replaces index.sh:79-83 + the dap.txt re-parse"}


def cpu_leg(wl, args):
    """C port of the reference algorithm on all host cores over a bounded sample, and
    the GPU rows / query of the same positions compared with it."""
    import numpy as np
    cores = os.cpu_count() or 1
    sample = min(wl.Lr, args.cpu_sample_rows)
    dap_np = wl.dap[:sample].cpu().numpy()
    dt, n_cpu, q_cpu, rows_cpu = cpu_port_run(dap_np, wl.rec_len, 0, wl.k, wl.n_docs, cores, order=wl.order)
    got = wl.owned_rows_between(0, sample)
    m = got[0].size
    same = (m == int((rows_cpu[0] < sample).sum()) and
            all(np.array_equal(g, w[:m]) for g, w in zip(got, rows_cpu)))
    if wl.membership:
        bits = wl.api.unpack_membership(wl.q_out[:sample - wl.k].cpu().numpy(), wl.n_docs)
        same = same and np.array_equal(bits, q_cpu[:sample - wl.k])
    else:
        same = same and np.array_equal(wl.q_out[:sample - wl.k].cpu().numpy(), q_cpu[:sample - wl.k])
    return {"value": sample / dt, "unit": "bp/s", "cores": cores, "kind": "port",
            "sample": f"first {sample} rows of the workload, one pass (index build + query); C port of the "
                      "reference algorithm (oracle/memo_oracle.c), one slice per core",
            "matches_gpu": bool(same), "rows_compared": int(m), "positions_compared": int(sample - wl.k)}


def extra_config(name, rec_len, C, membership, seed, k, dev, steps, peak, peak_src, tag):
    """A smaller BASELINE config measured after the headline one (N = 1 only): same step, plus
    the query timed alone with L2 flushed between launches (its index rows fit in L2)."""
    import torch
    wl = Workload(rec_len, C, k, membership, seed, 0, 1, dev, {})
    sync = torch.cuda.synchronize
    t = timed_run(wl, steps, 3, sync, dev.index or 0)
    wl.check_invariant()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    q_ms = 0.0
    for i in range(steps + 2):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wl.query()
        e1.record()
        sync()
        if i >= 2:
            q_ms += e0.elapsed_time(e1)
    q_ms /= steps
    r = rooflines(wl, t, peak, peak_src, wl.n_all, tag)
    bytes_q = r["roofline_query"]["algorithmic_bytes"]
    return {"config": name, "genomes": C + 1, "pivot_bp": rec_len, "membership": membership,
            "value": rec_len * steps / (t["total_ms"] * 1e-3), "unit": "bp/s",
            "ms_per_step": t["total_ms"] / steps, "index_ms": t["idx_ms"], "query_ms": t["qry_ms"],
            "index_rows": wl.n_all, "roofline": r["roofline"], "roofline_index_build": r["roofline_index_build"],
            "roofline_query": r["roofline_query"], "clocks": t["clocks"],
            "roofline_query_l2_flushed": {"query_ms": q_ms, "achieved": bytes_q / (q_ms * 1e-3) / 1e9,
                                          "frac": bytes_q / (q_ms * 1e-3) / 1e9 / peak,
                                          "note": "query timed alone, 512 MB written between launches"}}


GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
          138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
          83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
WG_KS = [15, 21, 31, 41, 51, 61, 71, 81, 91, 101]


def run_whole_genome(args, rank, world, local, dev, barrier):
    """BASELINE configs[4]: 94 genomes x the 24 GRCh38 primary records (3 088 269 832 bp), the
    concatenated pivot cut into one position range per GPU; index build + k sweep (15 .. 101,
    one fused launch).  1.15 TB of DAP does not fit the GPUs at once: every rank walks its range
    in pieces of one record and at most --wg-piece-rows rows, generated on the device (untimed),
    built and queried with the piece resident (timed with CUDA events); times add up per rank,
    the slowest rank decides."""
    import torch
    import torch.distributed as dist
    from memo_b200 import api, shard, _lib
    C, ks = args.cols, WG_KS
    n_docs = C + 1
    lens = [max(1000, int(n * args.wg_scale)) for n in GRCH38]     # (--wg-scale < 1: smoke runs only)
    total = sum(lens)
    g_lo, g_hi = shard.shard_range(total, world, rank)
    pieces, acc = [], 0
    for rid, n in enumerate(lens):
        a, b = max(g_lo, acc), min(g_hi, acc + n)
        while a < b:
            e = min(b, a + args.wg_piece_rows)
            pieces.append((rid, n, a - acc, e - acc))
            a = e
        acc += n
    lib = _lib.load()
    peak, peak_src = load_peaks()
    t_idx = t_qry = t_kern = 0.0
    n_rows = bytes_idx = bytes_q = 0
    bp = 0
    for rid, rec_len, lo, hi in pieces:
        wl = Workload(rec_len, C, 31, False, SEED0 + 4 + rid, 0, 1, dev, {}, span=(lo, hi))
        n, o = wl.n_all, wl.out
        sweep = lambda: api.query_sweep(o[0][:n], o[1][:n], o[2][:n], lo, hi, ks, n_docs, False,
                                        workspace=wl.q_ws, check=False)
        wl.build()
        res = sweep()                                     # warm-up + the result that is checked
        torch.cuda.synchronize()
        off = wl.lo - wl.buf_lo
        for i, k in enumerate(ks):                        # conservation == 1 + #{MS >= k}, every k
            for a in range(0, wl.Lr, 4_000_000):
                b = min(a + 4_000_000, wl.Lr)
                want = (1 + (wl.dap[off + a:off + b] >= k).sum(dim=1)).to(torch.uint8)
                assert torch.equal(res[i, a:b], want), f"sweep result violates the invariant at k={k}"
        del res
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        lib.memo_profile_enable(1)
        reps = args.wg_reps
        for _ in range(reps):
            ev[0].record()
            wl.build()
            ev[1].record()
            r = sweep()
            ev[2].record()
            torch.cuda.synchronize()
            t_idx += ev[0].elapsed_time(ev[1]) / reps
            t_qry += ev[1].elapsed_time(ev[2]) / reps
            del r
        lib.memo_profile_enable(0)
        k_ms, k_n = ctypes.c_double(0.0), ctypes.c_int32(0)
        _lib.check(lib.memo_profile_collect(ctypes.byref(k_ms), ctypes.byref(k_n)), "memo_profile_collect")
        t_kern += k_ms.value / max(k_n.value, 1)
        owned = int(wl.seg_out_end[0].item())
        n_rows += owned
        bytes_idx += 4.0 * wl.Lr * C + 12.0 * owned
        bytes_q += 12.0 * n + float(len(ks)) * wl.Lr      # 12 n_in + K W
        bp += wl.Lr
        del wl, o
        torch.cuda.empty_cache()
    stats = torch.tensor([t_idx, t_qry, t_kern, t_idx + t_qry, float(n_rows), bytes_idx, bytes_q, float(bp)],
                         dtype=torch.float64, device=dev)
    mx, sm = stats.clone(), stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        t_i, t_q, t_k, t_all = (mx[i].item() for i in range(4))
        line = {
            "metric": "pivot bp/s (conservation index build + k-mer query sweep, k = 15 .. 101)",
            "value": sm[7].item() / (t_all * 1e-3), "unit": "bp/s", "n_gpus": world,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: whole genome (24 GRCh38 primary records, {total} bp "
                                   f"pivot{'' if args.wg_scale == 1 else ', lengths scaled by %g' % args.wg_scale}) x {n_docs} genomes, conservation index build + multi-k query sweep "
                                   f"k = {ks} (one fused launch per piece), position ranges x{world}",
                       "genomes": n_docs, "pivot_bp": total, "ks": ks, "pieces_rank0": len(pieces),
                       "piece_rows_max": args.wg_piece_rows, "reps_per_piece": args.wg_reps,
                       "timing": "per piece: device-resident DAP, CUDA events around build and sweep; generation "
                                 "untimed; per-rank sums, max over ranks"},
            "ms_total": t_all, "index_ms": t_i, "query_ms": t_q,
            "index_bp_per_s": sm[7].item() / (t_i * 1e-3), "query_bp_per_s": sm[7].item() / (t_q * 1e-3),
            "index_rows": int(sm[4].item()),
            # aggregate bytes of all ranks / slowest rank's time, against N x the per-GPU peak
            "roofline_index_build": {"bound": "hbm", "achieved": sm[5].item() / (t_i * 1e-3) / 1e9,
                                     "peak": peak * world, "unit": "GB/s",
                                     "frac": sm[5].item() / (t_i * 1e-3) / 1e9 / (peak * world),
                                     "algorithmic_bytes": sm[5].item(), "peak_source": peak_src},
            "roofline": {"kernel": "wide_kernel (streaming kernel of the build)", "bound": "hbm",
                         "achieved": sm[5].item() / (t_k * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": sm[5].item() / (t_k * 1e-3) / 1e9 / (peak * world), "traffic": None},
            "roofline_query": {"kernel": "query_planes_kernel, 10 k values per launch", "bound": "hbm",
                               "achieved": sm[6].item() / (t_q * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                               "frac": sm[6].item() / (t_q * 1e-3) / 1e9 / (peak * world),
                               "algorithmic_bytes": sm[6].item()},
        }
        print(json.dumps(line))


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="memo_b200")
    ap.add_argument("--rows", type=int, default=CHR1, help="pivot bp of the record (cut into --gpus position shards)")
    ap.add_argument("--cols", type=int, default=93, help="DAP columns (genomes - 1)")
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-rows", type=int, default=16_000_000, help="host-resident slice per GPU of the e2e leg")
    ap.add_argument("--e2e-text-rows", type=int, default=1_000_000, help="rows of the dap.txt slice of the e2e_text figure")
    ap.add_argument("--ref-sample-rows", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=8_000_000)
    ap.add_argument("--parity-rows", type=int, default=1_000_000, help="positions either side of a shard cut compared with the port")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] / configs[2] figures")
    ap.add_argument("--rows-per-tile", type=int, default=0)
    ap.add_argument("--emit-buf", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="index kernel variant (see memo_index_opts_t)")
    ap.add_argument("--membership", action="store_true",
                    help="membership index (-m: no --order) + membership query (BASELINE configs[2])")
    ap.add_argument("--env", action="append", default=[], help="KEY=VAL set before the library loads (tuning)")
    ap.add_argument("--wg", action="store_true",
                    help="BASELINE configs[4]: whole genome (24 GRCh38 records) x 94, index build + k sweep")
    ap.add_argument("--wg-piece-rows", type=int, default=64_000_000)
    ap.add_argument("--wg-reps", type=int, default=2)
    ap.add_argument("--wg-scale", type=float, default=1.0, help="scale the record lengths (smoke runs)")
    args = ap.parse_args()
    for kv in args.env:
        key, _, val = kv.partition("=")
        os.environ[key] = val
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from memo_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.wg:
        run_whole_genome(args, rank, world, local, dev, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    tuning = dict(rows_per_tile=args.rows_per_tile, emit_buf_records=args.emit_buf,
                  warps_per_cta=args.warps, ctas_per_sm=args.ctas_per_sm, stages=args.stages,
                  kernel_variant=args.variant)
    wl = Workload(args.rows, args.cols, args.k, args.membership, seed_of(args), rank, world, dev, tuning)
    t = timed_run(wl, args.steps, args.warmup, barrier, local)
    wl.check_invariant()                                 # correctness guard on the timed configuration
    n_owned = int(wl.seg_out_end[0].item())

    stats = torch.tensor([t["total_ms"], t["idx_ms"], t["qry_ms"], t["kern_ms"], float(n_owned),
                          float(t["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx, sm = stats.clone(), stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        tmax = {"total_ms": mx[0].item(), "idx_ms": mx[1].item(), "qry_ms": mx[2].item(), "kern_ms": mx[3].item()}
        n_owned_total, launches = int(sm[4].item()), int(sm[5].item())
    else:
        tmax, n_owned_total, launches = t, n_owned, t["launches"]

    parity = None
    if world > 1:
        ok, m = shard_parity(wl, args)
        parity = {"ok": ok, "positions_each_side": m, "cuts": world - 1}
    e2e = None if (args.no_e2e or args.membership) else e2e_leg(wl, args, barrier)
    e2e_text, lengths_job = None, None
    if rank == 0 and world == 1 and not args.no_e2e and not args.membership:
        e2e_text = e2e_text_leg(wl, args)
        import types                                       # the --lengths leg runs last (below), from this host copy
        lengths_job = (types.SimpleNamespace(C=wl.C, rec_len=wl.rec_len, order=wl.order),
                       wl.dap[:min(wl.Lr, args.e2e_text_rows)].cpu().numpy())
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_leg(wl, args)

    if rank == 0:
        peak, peak_src = load_peaks()
        tag = f"c{args.cols}_{'memb' if args.membership else 'cons'}_{wl.Lr}"
        # max-over-ranks times against one shard's bytes (near-equal shards)
        r = rooflines(wl, {**t, **tmax}, peak, peak_src, n_owned, tag)
        total_ms = tmax["total_ms"]
        line = {
            "metric": "pivot bp/s (%s index build + k-mer query)" % ("membership" if args.membership else "conservation"),
            "value": args.rows * args.steps / (total_ms * 1e-3), "unit": "bp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args, world),
            "index_bp_per_s": args.rows / (tmax["idx_ms"] * 1e-3),
            "query_bp_per_s": args.rows / (tmax["qry_ms"] * 1e-3),
            "index_ms": tmax["idx_ms"], "query_ms": tmax["qry_ms"],
            "index_rows": n_owned_total, "rho_cell": n_owned_total / (args.rows * args.cols),
            "index_rows_parked_rank0": t["parked_rows"],
            **r,
            "cpu_baseline": cpu, "e2e": e2e, "e2e_text": e2e_text,
            "shard_parity": None if parity is None else parity["ok"], "shard_parity_detail": parity,
            "gpu_launches": launches, "clocks": t["clocks"],
        }
    del wl
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_extras and args.rows == CHR1 and args.cols == 93:
        # (the 92.6 GB just released keep the memory system busy for a few hundred ms -- a 10-genome
        #  build started right away runs at 0.71 ms instead of 0.59, scripts/narrow_probe.py)
        time.sleep(1.0)
        ex_steps = max(5, min(args.steps, 20))
        line["extra_configs"] = [
            extra_config("BASELINE configs[1]: 10 genomes x 100 Mbp, conservation", 100_000_000, 9, False,
                         SEED0 + 1, args.k, dev, ex_steps, peak, peak_src, "c9_cons_100000000"),
            extra_config("BASELINE configs[2]: 94 haplotypes x 5 Mbp, membership (-m)", 5_000_000, 93, True,
                         SEED0 + 2, args.k, dev, ex_steps, peak, peak_src, "c93_memb_5000000"),
            extra_config("94 genomes x 10 Mbp, conservation (round-1 comparison shape)", 10_000_000, 93, False,
                         SEED0 + 3, args.k, dev, ex_steps, peak, peak_src, "c93_cons_10000000"),
        ]
    if lengths_job is not None:
        # an extra, measured after everything else so that nothing it does can touch the line's
        # other figures; a failure is reported in its key
        try:
            line["e2e_text"]["from_lengths_files"] = e2e_lengths_leg(*lengths_job)
        except Exception as exc:
            line["e2e_text"]["from_lengths_files"] = {"error": repr(exc)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
