"""The C restatement (CPU baseline port) against the pinned numpy oracle and goldens."""
import numpy as np

from oracle import c_oracle as co
from oracle import memo_oracle as mo


def test_c_oracle_fuzz_goldens(fuzz_golden):
    arrays, meta = fuzz_golden
    for case in meta:
        name = case["name"]
        recs = [tuple(r) for r in case["records"]]
        vals = arrays[f"{name}.vals"]
        hdrs = [h for h, _ in recs]
        for order, tag in ((True, "cons"), (False, "memb")):
            want = arrays[f"{name}.{tag}.rows"].astype(np.int64)
            want_rec = arrays[f"{name}.{tag}.rec"].astype(np.int64)
            r, s, e, c = co.index_build(vals, recs, order)
            assert np.array_equal(np.stack([s, e, c], 1), want), (name, tag)
            assert np.array_equal(r, want_rec)
            for q in case["queries"]:
                if q["membership"] != (not order):
                    continue
                m = want_rec == hdrs.index(q["rec"])
                out = co.query(want[m, 0], want[m, 1], want[m, 2], q["s"], q["e"], q["k"], q["n"],
                               q["membership"])
                assert np.array_equal(out, arrays[q["key"]]), q


def test_c_oracle_vs_numpy_synth():
    L, C = 40000, 9
    vals = mo.synth_dap(L, C, seed=5, dense=True)
    recs = [("a", 15000), ("b", 25000)]
    for order in (True, False):
        a = co.index_build(vals, recs, order)
        b = mo.index_build(vals, recs, order)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        m = b[0] == 1
        for memb in (False, True):
            if memb == order:
                continue
            q1 = co.query(b[1][m], b[2][m], b[3][m], 50, 20000, 31, C + 1, memb)
            q2 = mo.query(b[1][m], b[2][m], b[3][m], 50, 20000, 31, C + 1, memb)
            assert np.array_equal(q1, q2)


def test_c_oracle_halo_slices_match_for_valid_ms():
    """The multi-threaded CPU baseline cuts records into slices with a one-row
    halo; for valid matching statistics the concatenation equals the whole."""
    L, C = 30000, 5
    vals = mo.synth_dap(L, C, seed=9)
    recs = [("a", L)]
    whole = co.index_build(vals, recs, True)
    parts = []
    cuts = [0, 7000, 7001, 19000, L]
    for i in range(len(cuts) - 1):
        a, b = cuts[i], cuts[i + 1]
        lo = a - 1 if a else 0
        segs = co.make_segs(recs, b - a, pos_first=a, row0=a - lo, primed_first=(a == 0),
                            chr_end_last=(b == L))
        parts.append(co.index_build(vals[lo:b], recs, True, segs=segs))
    for j in range(4):
        assert np.array_equal(np.concatenate([p[j] for p in parts]), whole[j])


def test_c_synth_matches_numpy_synth():
    """bench.py's reference arm generates its sample with the C generator."""
    for dense in (False, True):
        a = mo.synth_dap(3_000_000, 7, 99, row0=123_456, rows=70_000, dense=dense)
        b = co.synth_dap(3_000_000, 7, 99, row0=123_456, rows=70_000, dense=dense)
        assert np.array_equal(a, b)
        a = mo.synth_dap(200_000, 5, 3, dense=dense)
        b = co.synth_dap(200_000, 5, 3, dense=dense, threads=3)
        assert np.array_equal(a, b)
