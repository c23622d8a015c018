#!/usr/bin/env python3
"""Host-only stages of the drop-in, timed where no GPU is needed (the build container):

  1. `--lengths` ingest: per-genome MONI files -> int32 DAP blocks
     (memo_b200.io.iter_lengths_columns on memo_lengths_block_parse), against the shell pipeline
     it replaces (index.sh:79 per file + `paste | nl` of :83, run on the same files) and against
     numpy's text parser (what the ingest used before);
  2. BED -> Parquet (src/parquet_compress_bed.py:16-39): memo_b200.parquet_compress_bed against the
     UNMODIFIED reference script, both as child processes on the same BED, tables compared.

Prints one JSON object.   python scripts/host_stage_bench.py [--cols 93 --rows 1000000]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402   (synthetic DAP / index rows for the inputs)
from oracle import memo_oracle as mo  # noqa: E402


def best(fn, reps):
    out = []
    for _ in range(reps):
        t = time.perf_counter()
        r = fn()
        out.append(time.perf_counter() - t)
    return min(out), r


def lengths_stage(d, cols, rows, reps):
    from memo_b200 import io
    vals = mo.synth_dap(rows, cols, seed=20240614).astype(np.int32)
    paths = []
    for j in range(cols):
        paths.append(os.path.join(d, f"g{j:03d}.w_rc.lengths"))
        with open(paths[-1], "w") as fh:                       # MONI layout: header + one line per record
            fh.write(">chrS\n" + " ".join(map(str, vals[:, j].tolist())) + "\n")
    size = sum(os.path.getsize(p) for p in paths)

    def ours():
        n = 0
        for b in io.iter_lengths_columns(paths):
            assert np.array_equal(b, vals[n:n + len(b)])
            n += len(b)
        return n

    def ours_untimed_check():
        n = 0
        for b in io.iter_lengths_columns(paths):
            n += len(b)
        return n

    ours()                                                     # correctness once, then time without the compare
    t_ours, n = best(ours_untimed_check, reps)
    res = {"files": cols, "rows": rows, "text_bytes": size, "host_threads": min(16, os.cpu_count() or 1, cols),
           "iter_lengths_columns": {"seconds": t_ours, "text_gbs": size / t_ours / 1e9, "bp_per_s": n / t_ours}}

    def numpy_parser():                                        # the parser the ingest used before (one column)
        data = open(paths[0], "rb").read().split(b"\n", 1)[1]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            return np.fromstring(data.decode("ascii"), dtype=np.int64, sep=" ").size

    t_np, _ = best(numpy_parser, 1)
    res["numpy_fromstring_one_file"] = {"seconds": t_np, "text_gbs": os.path.getsize(paths[0]) / t_np / 1e9}

    # the reference's own route on the same files: index.sh:79 per file, then paste | nl (index.sh:83)
    def shell():
        for p in paths:
            subprocess.run(f"cat {p} | grep -v '^>' | tr ' ' '\\n' | grep . > {p}.vert", shell=True, check=True)
        subprocess.run("paste -d ' ' " + " ".join(p + ".vert" for p in paths) + f" | nl -v0 -w1 -s' ' > {d}/dap.txt",
                       shell=True, check=True)
        return os.path.getsize(os.path.join(d, "dap.txt"))

    t_sh, dap_bytes = best(shell, 1)
    res["index_sh_79_83_shell_pipeline"] = {"seconds": t_sh, "text_gbs": size / t_sh / 1e9, "bp_per_s": rows / t_sh,
                                            "dap_txt_bytes": dap_bytes,
                                            "note": "only produces dap.txt, which dap_to_bed.py then parses again"}
    return res


def parquet_stage(d, ref, reps):
    import pyarrow.parquet as pq
    recs = [("chr1", 600_000), ("chr2", 400_000)]
    vals = np.concatenate([mo.synth_dap(n, 93, seed=3 + i) for i, (_, n) in enumerate(recs)]).astype(np.int64)
    rec, s, e, c = co.index_build(vals, recs, True)
    bed = os.path.join(d, "idx.bed")
    with open(bed, "w") as fh:
        for rep in range(8):                                   # 16 records
            names = np.array([f"{h}_{rep}" for h, _ in recs])[rec]
            fh.write("".join(f"{a}\t{b}\t{x}\t{y}\n" for a, b, x, y in zip(names, s, e, c)))
    res = {"bed_bytes": os.path.getsize(bed), "rows": 8 * len(s)}
    cmds = {"memo_b200": [sys.executable, "-m", "memo_b200.parquet_compress_bed", "-f", bed, "-o", os.path.join(d, "ours.parquet")]}
    script = os.path.join(ref, "src", "parquet_compress_bed.py")
    if os.path.isfile(script):
        cmds["reference"] = [sys.executable, script, "-f", bed, "-o", os.path.join(d, "ref.parquet")]
    times = {k: [] for k in cmds}
    for _ in range(reps):                                      # alternating, so that both see the same machine
        for k, cmd in cmds.items():
            t = time.perf_counter()
            subprocess.run(cmd, check=True, capture_output=True, cwd=ROOT)
            times[k].append(time.perf_counter() - t)
    for k, v in times.items():
        res[k] = {"seconds_min": min(v), "seconds_all": [round(x, 2) for x in v]}
    if "reference" in cmds:
        a, b = pq.read_table(os.path.join(d, "ref.parquet")), pq.read_table(os.path.join(d, "ours.parquet"))
        res["tables_equal"] = bool(a.equals(b) and a.schema.equals(b.schema, check_metadata=True))
        res["query_read"] = query_read_stage(d, ref, reps)
    return res


def query_read_stage(d, ref, reps):
    """memo_query.py:19-36 (filter_pq: two predicate scans of the whole Parquet) on the reference's
    file against io.read_index_rows (row groups pruned by their statistics) on ours: a 300 kbp
    window of one of the 16 records, k = 31."""
    import importlib.util
    from memo_b200 import io
    spec = importlib.util.spec_from_file_location("ref_memo_query", os.path.join(ref, "src", "memo_query.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rec, qs, qe, k = "chr1_3", 100_000, 400_000, 31
    t_ref, rows = best(lambda: mod.filter_pq(os.path.join(d, "ref.parquet"), rec, qs, qe + k), reps)
    stats = {}
    t_us, cols = best(lambda: io.read_index_rows(os.path.join(d, "ours.parquet"), rec, qs, qe + k, stats), reps)
    live = rows[rows[:, 0] > qs]                                  # the rows of the live predicate
    same = bool(np.array_equal(live[:, 0], cols[0]) and np.array_equal(live[:, 1], cols[1]) and
                np.array_equal(live[:, 2], cols[2]))
    return {"window": f"{rec}:{qs}-{qe}", "k": k, "reference_filter_pq_seconds": t_ref, "rows_reference": int(len(rows)),
            "read_index_rows_seconds": t_us, "rows": int(len(cols[0])), "row_groups": stats, "live_rows_equal": same}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("MEMO_REFERENCE", "/root/reference"))
    ap.add_argument("--cols", type=int, default=93)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        out = {"host": {"cpus": os.cpu_count()}, "lengths_ingest": lengths_stage(d, args.cols, args.rows, args.reps)}
    with tempfile.TemporaryDirectory() as d:
        out["parquet_stage"] = parquet_stage(d, args.ref, args.reps)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
