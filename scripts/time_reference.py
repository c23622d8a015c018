#!/usr/bin/env python3
"""Times the UNMODIFIED reference scripts (dap_to_bed.py, parquet_compress_bed.py, memo_query.py)
as child processes on text slices of the bench workloads: figure A = as
shipped (one process, one core), figure B = one dap_to_bed.py per slice on all cores.  Runs only
where the reference tree exists (the build container); prints one JSON object.

    python scripts/time_reference.py [--cols 93 --rows 100000] [--ref /root/reference]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402  (generates the slice)


def write_case(d, name, vals, rec_len):
    fai, dap = os.path.join(d, f"{name}.fa.fai"), os.path.join(d, f"{name}.dap.txt")
    with open(fai, "w") as fh:
        fh.write(f"chrS\t{rec_len}\t6\t{rec_len}\t{rec_len + 1}\n")
    with open(dap, "w") as fh:
        for i, row in enumerate(vals):
            fh.write(f"{i} " + " ".join(map(str, row)) + "\n")
    return fai, dap


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("MEMO_REFERENCE", "/root/reference"))
    ap.add_argument("--cols", type=int, default=93)
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=20240614)
    ap.add_argument("--k", type=int, default=31)
    args = ap.parse_args()
    src = os.path.join(args.ref, "src")
    cores = os.cpu_count() or 1
    C, L = args.cols, args.rows
    out = {"cols": C, "rows": L, "cores": cores, "python": sys.version.split()[0]}
    with tempfile.TemporaryDirectory() as d:
        vals = co.synth_dap(L, C, args.seed, threads=cores)           # one record of exactly L rows
        fai, dap = write_case(d, "whole", vals, L)
        bed = os.path.join(d, "whole.bed")
        t0 = time.perf_counter()
        with open(bed, "wb") as fh:
            subprocess.run([sys.executable, os.path.join(src, "dap_to_bed.py"), "--mem", "--order", "--overlap",
                            "--fai", fai, "--dap", dap], stdout=fh, check=True)
        t_idx = time.perf_counter() - t0
        pqf = os.path.join(d, "whole.parquet")
        t0 = time.perf_counter()
        subprocess.run([sys.executable, os.path.join(src, "parquet_compress_bed.py"), "-f", bed, "-o", pqf],
                       stdout=subprocess.DEVNULL, check=True)
        t_pq = time.perf_counter() - t0
        t0 = time.perf_counter()
        subprocess.run([sys.executable, os.path.join(src, "memo_query.py"), "-b", pqf, "-r", f"chrS:0-{L}",
                        "-k", str(args.k), "-n", str(C + 1), "-o", os.path.join(d, "q.txt")],
                       stdout=subprocess.DEVNULL, check=True)
        t_q = time.perf_counter() - t0
        out["figure_A_one_core"] = {
            "dap_to_bed_s": t_idx, "index_bp_per_s": L / t_idx, "parquet_compress_bed_s": t_pq,
            "memo_query_s": t_q, "query_bp_per_s": L / t_q,
            "index_plus_query_bp_per_s": L / (t_idx + t_pq + t_q)}
        # figure B: one dap_to_bed.py per slice (its own record: state resets per record, :129)
        per = L // cores
        cases = []
        for i in range(cores):
            sl = vals[i * per:(i + 1) * per]
            cases.append(write_case(d, f"s{i}", sl, len(sl)))

        def run(i):
            with open(os.path.join(d, f"s{i}.bed"), "wb") as fh:
                subprocess.run([sys.executable, os.path.join(src, "dap_to_bed.py"), "--mem", "--order", "--overlap",
                                "--fai", cases[i][0], "--dap", cases[i][1]], stdout=fh, check=True)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(run, range(cores)))
        t_all = time.perf_counter() - t0
        out["figure_B_all_cores"] = {"dap_to_bed_s": t_all, "index_bp_per_s": per * cores / t_all, "processes": cores}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
