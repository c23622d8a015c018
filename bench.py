#!/usr/bin/env python3
"""bench.py -- MEMO hot path on B200: pivot bp/s for conservation index build +
k=31 window query (BASELINE.json configs[1]: 10 genomes x 100 Mbp, synthetic
HPRC-shaped DAP), with the HBM roofline of the dominant kernel, an end-to-end
leg through the host-buffer API, and the CPU port timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm (oracle port, all host cores)

One "step" = one pass of the hot path over the workload: index build of the
rank's DAP shard (device resident) followed by the k-mer conservation query over
the shard's window on the freshly built index rows.  value = total pivot bp over
all ranks / max-over-ranks device time.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240611 + 1            # seed = 20240611 + config index
KH = 128                       # right halo rows kept for queries (k <= 129)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


CAPTURED = {9: 100_000_000, 93: 10_000_000}     # pivot bp of the workloads profiles/traffic.json was captured on


def ncu_traffic(cols, whole=False, rows=None):
    """dram bytes (read + write) per launch of the streaming kernel (or of the whole
    index build) of the default workload, from the committed `ncu --set full` capture
    (profiles/traffic.json, scripts/ncu_traffic.py), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if rows is not None and CAPTURED.get(cols) != rows:
        return None                                  # per launch of a different workload
    try:
        data = json.load(open(path))
        if whole:
            return data.get(f"index_build_c{cols}")
        for k in data.get(f"index_build_c{cols}_kernels", []):
            if "narrow_kernel" in k["kernel"] or "wide_kernel" in k["kernel"]:
                return k["dram_bytes"]
    except Exception:
        pass
    return None


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
def cpu_port_run(dap_np, rec_len, row0, k, n_docs, threads):
    """Index build (conservation) + k-mer query over the rows of `dap_np` (a
    slice [row0, row0+n) of one record), cut into `threads` slices with a
    one-row halo (exact for matching statistics).  Returns (seconds, n_out)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle as co

    n, C = dap_np.shape
    recs = [("chrS", rec_len)]
    cuts = [row0 + (n * i) // threads for i in range(threads + 1)]

    def work(i):
        a, b = cuts[i], cuts[i + 1]
        lo = a - 1 if a > row0 else a
        segs = co.make_segs(recs, b - a, pos_first=a, row0=a - lo, primed_first=(a == 0),
                            chr_end_last=(b == rec_len))
        cap = int((b - a) * C * 0.05) + 4 * C
        r = co.index_build(dap_np[lo - row0:b - row0], recs, True, segs=segs, cap=cap)
        return r

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, range(threads)))
    f1 = np.concatenate([p[1] for p in parts])
    f2 = np.concatenate([p[2] for p in parts])
    f3 = np.concatenate([p[3] for p in parts])
    qcuts = cuts

    def qwork(i):
        return co.query(f1, f2, f3, qcuts[i], qcuts[i + 1], k, n_docs, False)

    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(qwork, range(threads)))
    dt = time.perf_counter() - t0
    return dt, int(f1.size), np.concatenate(outs), (f1, f2, f3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import memo_oracle as mo
    C, k = args.cols, args.k
    cores = os.cpu_count() or 1
    sample = min(args.rows, args.ref_sample_rows)
    dap = mo.synth_dap(args.rows, C, SEED, row0=0, rows=sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt, n_out, _, _ = cpu_port_run(dap, args.rows, 0, k, C + 1, cores)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        "impl": "reference", "metric": "pivot bp/s (conservation index build + k-mer query)",
        "value": value, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "bp/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} rows of the workload per step; C port of the "
                                   "reference algorithm (oracle/memo_oracle.c), one slice per core"},
        "e2e": {"value": value, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    memb = getattr(args, "membership", False)
    if args.cols == 9 and args.rows == 100_000_000 and not memb:
        tag = "BASELINE configs[1]"
    elif memb:
        tag = "BASELINE configs[2]-shaped: membership index (-m) and per-genome presence bitmaps"
    elif args.cols == 93:
        tag = "BASELINE configs[3]-shaped shard: 94 genomes"
    else:
        tag = "non-default shape"
    what = "membership index build + membership" if memb else "conservation index build +"
    return {"workload": f"synthetic HPRC-shaped DAP, {args.cols + 1} genomes x {args.rows} bp pivot "
                        f"per GPU ({tag}), {what} k={args.k} "
                        "window query over the whole shard",
            "genomes": args.cols + 1, "pivot_bp_per_gpu": args.rows, "k": args.k,
            "partition": f"position ranges x{n_gpus}, 1-row left halo, {KH}-row right halo",
            "l2": f"inputs ({args.rows * args.cols * 4 / 1e9:.2f} GB DAP per GPU) are larger than L2; no flush needed",
            "seed": SEED}


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="memo_b200")
    ap.add_argument("--rows", type=int, default=100_000_000, help="pivot bp per GPU")
    ap.add_argument("--cols", type=int, default=9, help="DAP columns (genomes - 1)")
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-sample-rows", type=int, default=8_000_000)
    ap.add_argument("--cpu-sample-rows", type=int, default=20_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rows-per-tile", type=int, default=0)
    ap.add_argument("--emit-buf", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="index kernel variant (1 = strip kernel for narrow rows)")
    ap.add_argument("--membership", action="store_true",
                    help="membership index (-m: no --order) + membership query (BASELINE configs[2])")
    ap.add_argument("--env", action="append", default=[], help="KEY=VAL set before the library loads (tuning)")
    args = ap.parse_args()
    for kv in args.env:
        key, _, val = kv.partition("=")
        os.environ[key] = val
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from memo_b200 import api, host, _lib

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

    C, k, Lr = args.cols, args.k, args.rows
    n_docs = C + 1
    rec_len = Lr * world                      # one record, position-range sharded (weak scaling)
    lo, hi = rank * Lr, (rank + 1) * Lr
    buf_lo = lo - 1 if rank > 0 else lo       # 1-row left halo
    buf_hi = min(hi + KH, rec_len)            # right halo for the query (rows owned by rank+1)
    dap = api.synth_dap(rec_len, C, SEED, row0=buf_lo, rows=buf_hi - buf_lo, device=dev)
    segs = [api.Segment(row_begin=lo - buf_lo, n_rows=Lr, pos0=lo, rec_len=rec_len, rec_id=0,
                        flags=(api.MEMO_SEG_PRIMED if rank == 0 else 0) |
                              (api.MEMO_SEG_CHR_END if rank == world - 1 else 0))]
    if buf_hi > hi:
        segs.append(api.Segment(row_begin=hi - buf_lo, n_rows=buf_hi - hi, pos0=hi, rec_len=rec_len,
                                rec_id=0, flags=0))
    tuning = dict(rows_per_tile=args.rows_per_tile, emit_buf_records=args.emit_buf,
                  warps_per_cta=args.warps, ctas_per_sm=args.ctas_per_sm, stages=args.stages,
                  kernel_variant=args.variant)
    builder = api.IndexBuilder(dev)
    seg_out_end = torch.zeros(len(segs), dtype=torch.int64, device=dev)
    # size the outputs with a counting run
    order = not args.membership
    builder.launch(dap, C, segs, order, None, seg_out_end, **tuning)
    n_all, irregular, _ = builder.result()
    assert not irregular, "synthetic DAP must be valid matching statistics"
    out = tuple(torch.empty(n_all + 16, dtype=torch.int32, device=dev) for _ in range(3))
    q_out = torch.empty((Lr, (n_docs + 31) // 32), dtype=torch.int32, device=dev) if args.membership \
        else torch.empty(Lr, dtype=torch.uint8, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    q_status = torch.zeros(1, dtype=torch.int32, device=dev)
    q_ws = torch.empty(max(_lib.load().memo_query_workspace_bytes(Lr), 1), dtype=torch.uint8, device=dev)

    def step(ev=None):
        # nothing in here waits for the device: the row count is known from the sizing
        # run (and re-checked after the timed region), the query reads the fresh rows
        if ev:
            ev[0].record()
        builder.launch(dap, C, segs, order, out, seg_out_end, **tuning)
        if ev:
            ev[1].record()
        if world > 1:                                 # ordered write offsets: gather the counts
            dist.all_gather_into_tensor(counts, seg_out_end[:1])
        if ev:
            ev[2].record()
        if args.membership:
            api.query_membership(out[0][:n_all], out[1][:n_all], out[2][:n_all], lo, hi, k, n_docs,
                                 out=q_out, check=False, status=q_status, workspace=q_ws)
        else:
            api.query_conservation(out[0][:n_all], out[1][:n_all], out[2][:n_all], lo, hi, k, n_docs,
                                   out=q_out, check=False, status=q_status, workspace=q_ws)
        if ev:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    lib = _lib.load()
    lib.memo_profile_enable(1)          # CUDA events around the streaming kernel of every build
    t_begin.record()
    for i in range(args.steps):
        step(evs[i])
    t_end.record()
    barrier()
    n_out, irregular, replays = builder.result()
    assert n_out == n_all and not irregular and int(q_status.item()) == 0
    lib.memo_profile_enable(0)
    import ctypes
    k_ms, k_n = ctypes.c_double(0.0), ctypes.c_int32(0)
    _lib.check(lib.memo_profile_collect(ctypes.byref(k_ms), ctypes.byref(k_n)), "memo_profile_collect")
    kern_ms = k_ms.value / max(k_n.value, 1)
    clocks = sampler.stop()
    total_ms = t_begin.elapsed_time(t_end)
    idx_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    qry_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    n_owned = int(seg_out_end[0].item())

    # correctness guard on the timed configuration: SURVEY 0.2 invariant
    if args.membership:
        # membership[p] = [1, MS_1[p] >= k, ...]: compare a slice bit by bit
        n_chk = min(Lr, 2_000_000)
        bits = api.unpack_membership(q_out[:n_chk].cpu().numpy(), n_docs)
        want = (dap[lo - buf_lo:lo - buf_lo + n_chk] >= k).cpu().numpy().astype(np.uint8)
        assert bits[:, 0].all() and np.array_equal(bits[:, 1:], want), "membership != [1, MS >= k]"
        del bits, want
    else:
        step_rows = 8_000_000                        # (in slices: the comparison needs temporaries)
        for a in range(0, Lr, step_rows):
            b = min(a + step_rows, Lr)
            want = (1 + (dap[lo - buf_lo + a:lo - buf_lo + b] >= k).sum(dim=1)).to(torch.uint8)
            assert torch.equal(q_out[a:b], want), "query result violates conservation == 1 + #{MS >= k}"
            del want

    stats = torch.tensor([total_ms, idx_ms, qry_ms, float(n_owned), kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, idx_ms, qry_ms, kern_ms = mx[0].item(), mx[1].item(), mx[2].item(), mx[4].item()
        n_owned_total = int(sm[3].item())
    else:
        n_owned_total = n_owned

    # ---------------- end-to-end leg: host buffers through the host API
    e2e = None
    if not args.no_e2e and not args.membership:
        host_dap = torch.empty(tuple(dap.shape), dtype=torch.int32, pin_memory=True)
        host_dap.copy_(dap)
        h2d = d2h = 0
        # one untimed pass: pinned result blocks and device buffers get allocated here
        rows_ = host.build_index(host_dap, None, True, device=dev, segs=segs, raw=True, **tuning)
        q_ = host.query(rows_.start, rows_.end, rows_.order, lo, hi, k, n_docs, False, device=dev,
                        raw=True, trusted=True)
        del rows_, q_
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rows_ = q_ = None                  # drop the previous step's results (their pinned blocks recycle)
            rows_ = host.build_index(host_dap, None, True, device=dev, segs=segs, raw=True, **tuning)
            q_ = host.query(rows_.start, rows_.end, rows_.order, lo, hi, k, n_docs, False, device=dev,
                            raw=True, trusted=True)
            h2d += host_dap.numel() * 4 + 12 * rows_.n
            d2h += 12 * rows_.n + q_.size
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ok = (rows_.n == n_all and
                  np.array_equal(rows_.start, out[0][:n_all].cpu().numpy()) and
                  np.array_equal(rows_.end, out[1][:n_all].cpu().numpy().view(np.uint32)) and
                  np.array_equal(rows_.order, out[2][:n_all].cpu().numpy()) and
                  np.array_equal(q_, q_out.cpu().numpy()))
        assert e2e_ok, "host-buffer path disagrees with the device-resident path"
        e2e = {"value": Lr * world * args.e2e_steps / t.item(), "unit": "bp/s",
               "h2d_bytes_per_step": h2d // args.e2e_steps, "d2h_bytes_per_step": d2h // args.e2e_steps,
               "steps": args.e2e_steps,
               "ms_per_step": 1e3 * t.item() / args.e2e_steps,
               "note": "host.build_index + host.query: pinned host DAP streamed to the device in 256 MB "
                       "chunks overlapped with the build, index rows (12 B each) and the query result "
                       "(1 B per bp) copied back to host; wall clock"}
        del host_dap

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.membership:
        cores = os.cpu_count() or 1
        sample = min(Lr, args.cpu_sample_rows)
        dap_np = dap[:sample].cpu().numpy()
        dt, n_cpu, q_cpu, rows_cpu = cpu_port_run(dap_np, rec_len, 0, k, n_docs, cores)
        # same sample, same answer: the GPU result must equal the CPU port's
        m = int((out[0][:n_out] < sample).sum().item())
        same = (np.array_equal(out[0][:m].cpu().numpy(), rows_cpu[0][:m]) and
                np.array_equal(out[1][:m].cpu().numpy().view(np.uint32), rows_cpu[1][:m]) and
                np.array_equal(out[2][:m].cpu().numpy(), rows_cpu[2][:m]) and
                np.array_equal(q_out[:sample - k].cpu().numpy(), q_cpu[:sample - k]))
        cpu = {"value": sample / dt, "unit": "bp/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} rows of the workload, one pass; C port of the reference "
                         "algorithm (oracle/memo_oracle.c), one slice per core",
               "matches_gpu": bool(same)}

    if rank == 0:
        peak, peak_src = load_peaks()
        # algorithmic bytes of one build: DAP read once + index rows written once
        bytes_idx = 4.0 * Lr * C + 12.0 * n_all
        ach = bytes_idx / (kern_ms * 1e-3) / 1e9          # dominant kernel: the streaming kernel
        ach_build = bytes_idx / (idx_ms * 1e-3) / 1e9     # whole memo_index_build (stream + scan + gather)
        n_q_rows = n_all
        bytes_q = 12.0 * n_q_rows + (4.0 * ((n_docs + 31) // 32) if args.membership else 1.0) * Lr
        ach_q = bytes_q / (qry_ms * 1e-3) / 1e9
        value = Lr * world * args.steps / (total_ms * 1e-3)
        line = {
            "metric": "pivot bp/s (%s index build + k-mer query)" % ("membership" if args.membership else "conservation"),
            "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, world),
            "index_bp_per_s": Lr * world / (idx_ms * 1e-3),
            "query_bp_per_s": Lr * world / (qry_ms * 1e-3),
            "index_ms": idx_ms, "query_ms": qry_ms,
            "index_rows": n_owned_total, "rho_cell": n_owned_total / (Lr * world * C),
            "roofline": {"kernel": ("narrow_kernel" if C <= 16 else "wide_kernel") +
                                   " (streaming kernel of memo_index_build: DAP -> index rows; CUDA events "
                                   "around the launch, averaged over the timed steps)",
                         "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic(C, rows=args.rows), "peak_source": peak_src,
                         "algorithmic_bytes": bytes_idx, "kernel_ms": kern_ms},
            "roofline_index_build": {"kernel": "memo_index_build = streaming kernel + tile_scan + strip_gather",
                                     "bound": "hbm", "achieved": ach_build, "peak": peak, "unit": "GB/s",
                                     "frac": ach_build / peak, "algorithmic_bytes": bytes_idx,
                                     "traffic": ncu_traffic(C, whole=True, rows=args.rows)},
            "roofline_query": {"kernel": "query_stream_kernel (" + ("membership" if args.membership else "conservation") + ")", "bound": "hbm",
                               "achieved": ach_q, "peak": peak, "unit": "GB/s", "frac": ach_q / peak,
                               "algorithmic_bytes": bytes_q},
            "cpu_baseline": cpu, "e2e": e2e,
            # per step: index build = stream kernel + tile_scan + gather, query = one stream kernel
            "gpu_launches": 4 * args.steps, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
