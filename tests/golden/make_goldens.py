#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference scripts (``/root/reference/src/{dap_to_bed,parquet_compress_bed,
memo_query}.py``) in the build container.  The reference cannot travel to the
GPU box, so the inputs and the outputs it produced are committed here.

    python tests/golden/make_goldens.py            # rewrites the fixtures

Fixtures written:
  example.dap.txt / example.fai   SURVEY A.4 example: brute-force matching
      statistics of example/ref_1.fa against ref_2..5 (+ reverse complement,
      '$'-separated; mirrors index.sh:63-76 because MONI is not installed)
  example.cons.bed / example.memb.bed          dap_to_bed.py outputs
  example.queries.json                         memo_query.py outputs
  fuzz.npz + fuzz.json                         random DAPs (valid MS, arbitrary
      ints, zeros, ties, several records, partial records) with the reference's
      BED rows and query outputs
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MEMO_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
PY = sys.executable


def run(args, **kw):
    return subprocess.run(args, check=True, capture_output=True, text=True, **kw)


# ---------------------------------------------------------------- example DAP
def read_fasta(path):
    recs, name, seq = [], None, []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(seq)))
            name, seq = line[1:].split()[0], []
        elif line:
            seq.append(line.upper())
    recs.append((name, "".join(seq)))
    return recs


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def brute_ms(pivot, text):
    out = []
    for p in range(len(pivot)):
        best = 0
        for l in range(1, len(pivot) - p + 1):
            if pivot[p:p + l] in text:
                best = l
            else:
                break
        out.append(best)
    return out


def example_dap():
    ex = os.path.join(REF, "example")
    pivot = read_fasta(os.path.join(ex, "ref_1.fa"))
    assert len(pivot) == 1
    cols = []
    for i in range(2, 6):
        recs = read_fasta(os.path.join(ex, f"ref_{i}.fa"))
        text = "$".join([s for _, s in recs] + [revcomp(s) for _, s in recs]) + "$"
        cols.append(brute_ms(pivot[0][1], text))
    L = len(pivot[0][1])
    dap = "".join(f"{p} " + " ".join(str(c[p]) for c in cols) + "\n" for p in range(L))
    fai = f"{pivot[0][0]}\t{L}\t7\t{L}\t{L + 1}\n"
    return dap, fai


# ------------------------------------------------------------- reference runs
def ref_index(tmp, dap_txt, fai_txt, order):
    fai = os.path.join(tmp, "p.fa.fai")
    dap = os.path.join(tmp, "dap.txt")
    open(fai, "w").write(fai_txt)
    open(dap, "w").write(dap_txt)
    args = [PY, os.path.join(SRC, "dap_to_bed.py"), "--mem", "--overlap",
            "--fai", fai, "--dap", dap]
    if order:
        args.insert(3, "--order")
    return run(args).stdout


def ref_query(tmp, bed_txt, region, k, n, membership):
    bed = os.path.join(tmp, "i.bed")
    pq = os.path.join(tmp, "i.parquet")
    out = os.path.join(tmp, "q.txt")
    open(bed, "w").write(bed_txt)
    run([PY, os.path.join(SRC, "parquet_compress_bed.py"), "-f", bed, "-o", pq])
    args = [PY, os.path.join(SRC, "memo_query.py"), "-b", pq, "-r", region,
            "-k", str(k), "-n", str(n), "-o", out]
    if membership:
        args.insert(2, "-m")
    run(args)
    return open(out).read()


def sha(s):
    return hashlib.sha256(s.encode()).hexdigest()


# ------------------------------------------------------------------ fuzz DAPs
def valid_ms(rng, n, C, rec_len, p0=0, jump=0.05, maxlen=40):
    d = np.where(rng.random((n, C)) < jump, rng.integers(1, maxlen, (n, C)),
                 rng.integers(0, 3, (n, C)))
    p = (p0 + np.arange(n))[:, None]
    reach = np.maximum.accumulate(d + p, axis=0)
    return np.minimum(reach - p, rec_len - p)


def fuzz_cases(rng):
    cases = []
    # (name, records [(hdr,len)], pos, vals)
    for i in range(6):                                   # valid MS, 1 record
        C = int(rng.integers(1, 8)); n = int(rng.integers(2, 60))
        cases.append((f"valid1_{i}", [("chrA", n)], np.arange(n),
                      valid_ms(rng, n, C, n)))
    for i in range(6):                                   # valid MS, 3 records
        C = int(rng.integers(1, 12))
        lens = [int(x) for x in rng.integers(1, 40, 3)]
        vals = np.concatenate([valid_ms(rng, n, C, n) for n in lens])
        recs = [(f"c{j}", n) for j, n in enumerate(lens)]
        cases.append((f"valid3_{i}", recs, np.arange(sum(lens)), vals))
    for i in range(8):                                   # arbitrary ints
        C = int(rng.integers(1, 10))
        lens = [int(x) for x in rng.integers(1, 30, int(rng.integers(1, 4)))]
        n = sum(lens)
        hi = int(rng.choice([2, 4, 12, 60]))
        vals = rng.integers(0, hi, (n, C))
        recs = [(f"r{j}", m) for j, m in enumerate(lens)]
        cases.append((f"arb_{i}", recs, np.arange(n), vals))
    for i in range(3):                                   # all zeros / all ties
        C = int(rng.integers(2, 6)); n = int(rng.integers(3, 20))
        cases.append((f"zero_{i}", [("z", n)], np.arange(n),
                      np.zeros((n, C), dtype=np.int64)))
        cases.append((f"tie_{i}", [("t", n)], np.arange(n),
                      np.full((n, C), int(rng.integers(1, 9)))))
    for i in range(3):                                   # partial: DAP shorter than fai
        C = int(rng.integers(1, 6))
        lens = [20, 30]
        n = int(rng.integers(21, 45))
        vals = rng.integers(0, 15, (n, C))
        cases.append((f"partial_{i}", [("a", 20), ("b", 30)], np.arange(n), vals))
    for i in range(2):                                   # wide rows (C > 32, > 64)
        C = 40 + 60 * i; n = 25
        cases.append((f"wide_{i}", [("w", n)], np.arange(n),
                      valid_ms(rng, n, C, n, jump=0.1)))
    # a record with zero length in the middle of the fai
    C = 3
    vals = rng.integers(0, 9, (15, C))
    cases.append(("emptyrec", [("a", 7), ("e", 0), ("b", 8)], np.arange(15), vals))
    return cases


def bed_rows(txt):
    rows = [l.split("\t") for l in txt.splitlines()]
    names = [r[0] for r in rows]
    arr = np.array([[int(x) for x in r[1:]] for r in rows], dtype=np.int64).reshape(-1, 3)
    return names, arr


def main():
    tmp = tempfile.mkdtemp(prefix="memo_golden_")
    dap, fai = example_dap()
    open(os.path.join(HERE, "example.dap.txt"), "w").write(dap)
    open(os.path.join(HERE, "example.fai"), "w").write(fai)
    cons = ref_index(tmp, dap, fai, True)
    memb = ref_index(tmp, dap, fai, False)
    open(os.path.join(HERE, "example.cons.bed"), "w").write(cons)
    open(os.path.join(HERE, "example.memb.bed"), "w").write(memb)
    print("example dap  sha256", sha(dap))
    print("example cons sha256", sha(cons), len(cons.splitlines()), "rows")
    print("example memb sha256", sha(memb), len(memb.splitlines()), "rows")
    queries = []
    for (m, k, region) in [(False, 3, "ref_1:0-20"), (True, 3, "ref_1:0-20"),
                           (False, 31, "ref_1:0-20"), (True, 31, "ref_1:0-20"),
                           (False, 3, "ref_1:5-26"), (False, 3, "ref_1:20-30"),
                           (False, 1, "ref_1:0-26"), (False, 5, "ref_1:0-26"),
                           (False, 3, "nochr:0-5"), (True, 5, "ref_1:3-26"),
                           (True, 2, "ref_1:20-30")]:
        out = ref_query(tmp, memb if m else cons, region, k, 5, m)
        queries.append({"membership": m, "k": k, "region": region, "n": 5,
                        "out": out, "sha256": sha(out)})
    json.dump(queries, open(os.path.join(HERE, "example.queries.json"), "w"), indent=1)

    rng = np.random.default_rng(20240611)
    arrays, meta = {}, []
    for name, recs, pos, vals in fuzz_cases(rng):
        vals = np.asarray(vals, dtype=np.int64)
        dap_txt = "".join(f"{int(p)} " + " ".join(str(int(x)) for x in row) + "\n"
                          for p, row in zip(pos, vals))
        fai_txt = "".join(f"{h}\t{n}\t0\t60\t61\n" for h, n in recs)
        entry = {"name": name, "records": recs, "queries": []}
        arrays[f"{name}.vals"] = vals.astype(np.int32)
        for order in (True, False):
            bed = ref_index(tmp, dap_txt, fai_txt, order)
            names, arr = bed_rows(bed)
            tag = "cons" if order else "memb"
            hdr_to_idx = {}
            for j, (h, _) in enumerate(recs):
                hdr_to_idx.setdefault(h, j)
            arrays[f"{name}.{tag}.rec"] = np.array([hdr_to_idx[x] for x in names], dtype=np.int32)
            arrays[f"{name}.{tag}.rows"] = arr.astype(np.int32)
            entry[f"{tag}_sha256"] = sha(bed)
            if not arr.size:
                continue                      # parquet_compress_bed.py crashes on empty BED
            n_docs = vals.shape[1] + 1
            for qi in range(3):
                h, n = recs[int(rng.integers(0, len(recs)))]
                s = int(rng.integers(0, max(1, n)))
                e = int(rng.integers(s + 1, n + 6))
                k = int(rng.choice([1, 2, 3, 5, 8, 31]))
                out = ref_query(tmp, bed, f"{h}:{s}-{e}", k, n_docs, not order)
                key = f"{name}.{tag}.q{qi}"
                if order:
                    arrays[key] = np.array(out.split(), dtype=np.int32)
                else:
                    arrays[key] = np.array([l.split() for l in out.splitlines()],
                                           dtype=np.uint8).reshape(e - s, n_docs)
                entry["queries"].append({"key": key, "membership": not order, "rec": h,
                                         "s": s, "e": e, "k": k, "n": n_docs,
                                         "sha256": sha(out)})
        meta.append(entry)
    np.savez_compressed(os.path.join(HERE, "fuzz.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "fuzz.json"), "w"), indent=1)
    print("fuzz cases:", len(meta))


if __name__ == "__main__":
    main()
