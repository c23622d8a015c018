"""The oracle (numpy restatement) against outputs of the reference itself."""
import hashlib

import numpy as np
import pytest

from oracle import memo_oracle as mo


def _bed_arrays(txt):
    rows = [l.split("\t") for l in txt.splitlines()]
    return [r[0] for r in rows], np.array([[int(x) for x in r[1:]] for r in rows],
                                          dtype=np.int64).reshape(-1, 3)


def test_example_sha256_match_survey(example_golden):
    sha = lambda s: hashlib.sha256(s.encode()).hexdigest()
    assert sha(example_golden["dap_txt"]) == \
        "da11d96a5e7701b31f30d1212614ce38cf95e8c10ccdf2392c213c8e295b7857"
    assert sha(example_golden["cons_bed"]) == \
        "b3927029ab837bcf4129452940959acf245d606e0e211bee15a219df7e0eb87d"
    assert sha(example_golden["memb_bed"]) == \
        "6da7b5fac7a4cae5deed7f965aaed81f27cf4b20abb0df2ee886f0e0abb1dc08"


@pytest.mark.parametrize("order", [True, False])
@pytest.mark.parametrize("fn", ["index_build", "index_build_stream"])
def test_example_index(example_golden, order, fn):
    g = example_golden
    r, s, e, c = getattr(mo, fn)(g["vals"], g["records"], order, pos=g["pos"])
    assert mo.format_bed(g["records"], r, s, e, c) == (g["cons_bed"] if order else g["memb_bed"])


def test_example_queries(example_golden):
    g = example_golden
    for q in g["queries"]:
        _, arr = _bed_arrays(g["memb_bed"] if q["membership"] else g["cons_bed"])
        chrom, se = q["region"].split(":")
        s, e = map(int, se.split("-"))
        if chrom != "ref_1":
            arr = arr[:0]
        out = mo.query(arr[:, 0], arr[:, 1], arr[:, 2], s, e, q["k"], q["n"], q["membership"])
        txt = mo.format_membership(out) if q["membership"] else mo.format_conservation(out)
        assert txt == q["out"], q


def test_fuzz_index_and_queries(fuzz_golden):
    arrays, meta = fuzz_golden
    n_q = 0
    for case in meta:
        name = case["name"]
        recs = [tuple(r) for r in case["records"]]
        vals = arrays[f"{name}.vals"]
        for order, tag in ((True, "cons"), (False, "memb")):
            want_rec = arrays[f"{name}.{tag}.rec"]
            want = arrays[f"{name}.{tag}.rows"]
            for fn in (mo.index_build, mo.index_build_stream):
                r, s, e, c = fn(vals, recs, order)
                got = np.stack([s, e, c], axis=1)
                assert np.array_equal(got, want), (name, tag, fn.__name__)
                assert np.array_equal(r, want_rec), (name, tag)
            hdrs = [h for h, _ in recs]
            for q in case["queries"]:
                if q["membership"] != (not order):
                    continue
                ridx = hdrs.index(q["rec"])
                m = want_rec == ridx
                out = mo.query(want[m, 0], want[m, 1], want[m, 2], q["s"], q["e"], q["k"],
                               q["n"], q["membership"])
                assert np.array_equal(out, arrays[q["key"]]), q
                n_q += 1
    assert n_q > 100


def test_closed_form_vs_stream_random():
    rng = np.random.default_rng(7)
    for _ in range(200):
        C = int(rng.integers(1, 9))
        lens = [int(x) for x in rng.integers(1, 25, int(rng.integers(1, 4)))]
        n = int(rng.integers(1, sum(lens) + 1))
        vals = rng.integers(0, int(rng.choice([2, 5, 30])), (n, C))
        recs = [(f"r{i}", m) for i, m in enumerate(lens)]
        for order in (True, False):
            a = mo.index_build(vals, recs, order)
            b = mo.index_build_stream(vals, recs, order)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)


def test_valid_ms_query_invariant():
    """for valid matching statistics the pipeline collapses to
    conservation[p] = 1 + #{MS >= k}; membership[p] = [1, MS >= k]."""
    L, C, k = 6000, 9, 31
    ms = mo.synth_dap(L, C, seed=20240612)
    recs = [("chrS", L)]
    r, s, e, c = mo.index_build(ms, recs, True)
    cons = mo.query(s, e, c, 0, L, k, C + 1, False)
    assert np.array_equal(cons, 1 + (ms >= k).sum(axis=1))
    r, s, e, c = mo.index_build(ms, recs, False)
    memb = mo.query(s, e, c, 0, L, k, C + 1, True)
    want = np.concatenate([np.ones((L, 1), dtype=np.uint8), (ms >= k).astype(np.uint8)], axis=1)
    assert np.array_equal(memb, want)


def test_synth_is_valid_ms_and_chunk_invariant():
    L, C = 50000, 5
    full = mo.synth_dap(L, C, seed=3)
    assert (full[1:] >= full[:-1] - 1).all() and (full >= 1).all()
    part = mo.synth_dap(L, C, seed=3, row0=40000, rows=5000)
    assert np.array_equal(part, full[40000:45000])


def test_position_beyond_records_raises():
    with pytest.raises(Exception, match="beyond all intervals"):
        mo.index_build(np.ones((5, 2)), [("a", 3)], True)


def test_view_bins_known_answers():
    """`memo view` binning (src/plot_conservation.py:46-58): example of the walkthrough
    (example/README.md: 20 positions, 4 bins, 5 genomes) by hand."""
    vec = [5, 4, 5, 4, 3, 4, 4, 5, 5, 4, 5, 4, 3, 5, 3, 5, 4, 5, 4, 3]     # SURVEY A.4, k=3 ref_1:0-20
    comp = mo.view_bins(vec, 5, 4)
    assert comp.shape == (4, 6)
    assert comp[0].tolist() == [0, 0, 0, 0.2, 0.4, 0.4]
    assert comp[1].tolist() == [0, 0, 0, 0, 0.6, 0.4]
    assert comp[3].tolist() == [0, 0, 0, 0.2, 0.4, 0.4]
    assert np.allclose(comp.sum(axis=1), 1.0)
    import pytest
    with pytest.raises(ZeroDivisionError):
        mo.view_bins([1, 2], 5, 4)


def test_k_sweep_planes_follow_from_the_smaller_k():
    """A property of memo_query.py:42-63 a fused k sweep may rely on: a row paints
    [f2 - s - (k - 1), f1 - s) clipped to the window, so for k' > k every painted run grows to the
    left by k' - k and the only other change is rows that paint for the first time.  Per genome /
    order column:  painted(k') = dilate_left(painted(k), k' - k)  |  rows first non-empty at k'."""
    rng = np.random.default_rng(4)
    for case in range(30):
        L, C = int(rng.integers(200, 1200)), int(rng.integers(1, 9))
        if case % 3 == 0:
            vals = mo.synth_dap(L, C, seed=int(rng.integers(1, 1 << 30)), dense=True)
        else:
            vals = rng.integers(0, (60, 6)[case % 3 - 1], size=(L, C))
        _, f1, f2, f3 = mo.index_build(vals.astype(np.int64), [("r", L)], order=bool(case & 1))
        n_docs = C + 1
        qs = int(rng.integers(0, L // 2))
        qe = int(rng.integers(qs + 1, L + 30))
        W = qe - qs
        ks = sorted(set(int(x) for x in rng.integers(1, 120, 6)))
        prev = k_prev = None
        for k in ks:
            painted = 1 - mo.query(f1, f2, f3, qs, qe, k, n_docs, True)       # [W, n_docs]: 1 where a row painted
            if prev is not None:
                want = prev.copy()
                for j in range(1, k - k_prev + 1):                            # grow every run to the left
                    want[:-j] |= prev[j:]
                m = f1 > qs
                b = np.clip(f1[m] - qs, 0, W)
                a_prev, a_now = (np.clip(f2[m] - qs - (kk - 1), 0, W) for kk in (k_prev, k))
                first = (a_prev >= b) & (a_now < b)                           # rows that paint for the first time
                for a, e, j in zip(a_now[first], b[first], f3[m][first]):
                    want[a:e, j] = 1
                assert np.array_equal(want, painted), (case, k_prev, k)
            prev, k_prev = painted, k
