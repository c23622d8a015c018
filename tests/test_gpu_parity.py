"""GPU parity: the CUDA path (through the C ABI) against the oracle and the
reference-generated golden fixtures.  Bit-exact: everything here is integer."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import memo_oracle as mo  # noqa: E402  (checker only)


def _api():
    from memo_b200 import api
    return api


def gpu_index(vals, records, order, **kw):
    api = _api()
    dap = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.int32)).cuda()
    res = api.index_build(dap, records, order, **kw)
    return res, res.to_host()


def assert_index_equal(got, want, ctx=""):
    names = ("rec", "start", "end", "col")
    for g, w, n in zip(got, want, names):
        assert g.shape == w.shape, f"{ctx}: {n} count {g.shape} != {w.shape}"
        if not np.array_equal(g, w):
            i = int(np.flatnonzero(g != w)[0])
            raise AssertionError(f"{ctx}: {n} differs first at row {i}: got {g[i]} want {w[i]}")


def gpu_query(s, e, c, qs, qe, k, n_docs, membership):
    """Queries the bit-plane stream kernels take (uint8 conservation, membership) run
    through BOTH device paths (MEMO_QUERY_PLANES=0 selects the tile kernels that serve
    everything else) and must agree."""
    if (membership or n_docs <= 255) and "MEMO_QUERY_PLANES" not in os.environ:
        outs = []
        for mode in ("1", "0"):
            os.environ["MEMO_QUERY_PLANES"] = mode
            try:
                outs.append(gpu_query(s, e, c, qs, qe, k, n_docs, membership))
            finally:
                del os.environ["MEMO_QUERY_PLANES"]
        assert np.array_equal(outs[0][0], outs[1][0]), "bit-plane and tile query kernels disagree"
        return outs[0]
    api = _api()
    f1 = torch.from_numpy(s.astype(np.int32)).cuda()
    f2 = torch.from_numpy(e.astype(np.uint32).view(np.int32)).cuda()
    f3 = torch.from_numpy(c.astype(np.int32)).cuda()
    if membership:
        bits = api.query_membership(f1, f2, f3, qs, qe, k, n_docs)
        return api.unpack_membership(bits.cpu().numpy(), n_docs), bits
    out = api.query_conservation(f1, f2, f3, qs, qe, k, n_docs)
    return out.cpu().numpy().astype(np.int64), out


# ------------------------------------------------------------------ goldens
@pytest.mark.parametrize("order", [True, False])
def test_example_index_golden(example_golden, order):
    g = example_golden
    res, got = gpu_index(g["vals"], g["records"], order)
    assert mo.format_bed(g["records"], *got) == (g["cons_bed"] if order else g["memb_bed"])
    assert not res.irregular


def test_example_queries_golden(example_golden):
    api = _api()
    g = example_golden
    for q in g["queries"]:
        bed = g["memb_bed"] if q["membership"] else g["cons_bed"]
        arr = np.array([[int(x) for x in l.split("\t")[1:]] for l in bed.splitlines()], dtype=np.int64)
        chrom, se = q["region"].split(":")
        s, e = map(int, se.split("-"))
        if chrom != "ref_1":
            arr = arr[:0]
        out, dev = gpu_query(arr[:, 0], arr[:, 1], arr[:, 2], s, e, q["k"], q["n"], q["membership"])
        if q["membership"]:
            assert api.format_membership(dev, q["n"]).decode() == q["out"], q
        else:
            assert api.format_conservation(dev).decode() == q["out"], q


def test_fuzz_goldens(fuzz_golden):
    arrays, meta = fuzz_golden
    n_general = 0
    for case in meta:
        name = case["name"]
        recs = [tuple(r) for r in case["records"]]
        vals = arrays[f"{name}.vals"]
        hdrs = [h for h, _ in recs]
        for order, tag in ((True, "cons"), (False, "memb")):
            want_rec = arrays[f"{name}.{tag}.rec"].astype(np.int64)
            want = arrays[f"{name}.{tag}.rows"].astype(np.int64)
            for kw in ({}, {"rows_per_tile": 4, "emit_buf_records": 2}):
                res, got = gpu_index(vals, recs, order, **kw)
                assert_index_equal(got, (want_rec, want[:, 0], want[:, 1], want[:, 2]),
                                   f"{name}/{tag}/{kw}")
                n_general += res.general
            for q in case["queries"]:
                if q["membership"] != (not order):
                    continue
                m = want_rec == hdrs.index(q["rec"])
                out, _ = gpu_query(want[m, 0], want[m, 1], want[m, 2], q["s"], q["e"], q["k"],
                                   q["n"], q["membership"])
                assert np.array_equal(out, arrays[q["key"]]), q
    assert n_general > 0          # the arbitrary-integer cases must take the general path


# ------------------------------------------------------------------ synthetic
@pytest.mark.parametrize("C,dense", [(1, False), (9, False), (9, True), (93, False), (130, True)])
def test_synth_matches_oracle(C, dense):
    api = _api()
    L = 70000
    want = mo.synth_dap(L, C, seed=20240611 + C, dense=dense)
    got = api.synth_dap(L, C, seed=20240611 + C, dense=dense).cpu().numpy()
    assert np.array_equal(got, want)
    part = api.synth_dap(L, C, seed=20240611 + C, row0=33333, rows=12345, dense=dense).cpu().numpy()
    assert np.array_equal(part, want[33333:33333 + 12345])


GEOMS = [1, 2, 4, 5, 8, 9, 12, 13, 16, 24, 25, 32, 33, 48, 64, 65, 93, 96, 128, 129, 192, 256, 300]


@pytest.mark.parametrize("C", GEOMS)
@pytest.mark.parametrize("order", [True, False])
def test_index_valid_ms_all_geometries(C, order):
    L = 20000 if C <= 128 else 6000
    vals = mo.synth_dap(L, C, seed=C * 7 + 1, dense=(C % 2 == 0))
    recs = [("chrS", L)]
    want = mo.index_build(vals, recs, order)
    res, got = gpu_index(vals, recs, order)
    assert_index_equal(got, want, f"C={C} order={order}")
    assert not res.irregular and not res.general
    # tiny strips / tiny staging buffers: exercises look-back and the replay path
    res, got = gpu_index(vals[:3000], [("chrS", 3000)], order, rows_per_tile=8, emit_buf_records=4)
    assert_index_equal(got, mo.index_build(vals[:3000], [("chrS", 3000)], order), f"C={C} tiny")
    # the single-kernel strip build (index_wide2.cu: ordered in-place writes), default and tiny shapes
    res, got = gpu_index(vals, recs, order, kernel_variant=3)
    assert_index_equal(got, want, f"C={C} order={order} single-kernel")
    assert not res.irregular and not res.general
    res, got = gpu_index(vals[:3000], [("chrS", 3000)], order, kernel_variant=3, rows_per_tile=8, emit_buf_records=4)
    assert_index_equal(got, mo.index_build(vals[:3000], [("chrS", 3000)], order), f"C={C} single-kernel tiny")


@pytest.mark.parametrize("C", [3, 9, 20, 40, 93, 150])
@pytest.mark.parametrize("order", [True, False])
def test_index_arbitrary_ints_general_path(C, order):
    rng = np.random.default_rng(C)
    lens = [700, 1, 2500, 1333]
    L = sum(lens)
    vals = rng.integers(0, 50, (L, C))
    vals[rng.random((L, C)) < 0.3] = 0
    recs = [(f"r{i}", n) for i, n in enumerate(lens)]
    want = mo.index_build(vals, recs, order)
    res, got = gpu_index(vals, recs, order, rows_per_tile=64)
    assert res.irregular and res.general
    assert_index_equal(got, want, f"C={C} order={order}")


@pytest.mark.parametrize("order", [True, False])
def test_index_multi_record_and_partial(order):
    rng = np.random.default_rng(5)
    C = 9
    lens = [5000, 1, 1, 12000, 300, 7777]
    recs = [(f"c{i}", n) for i, n in enumerate(lens)] + [("unused", 1000)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=100 + i, dense=(i % 2 == 0))
                           for i, n in enumerate(lens)])
    vals = vals[:-100]                      # DAP ends inside the last record it touches
    want = mo.index_build(vals, recs, order)
    res, got = gpu_index(vals, recs, order)
    assert_index_equal(got, want, "multi")
    assert not res.general


@pytest.mark.parametrize("C", [1, 2, 3, 4, 6, 7, 8, 9, 12, 15, 16])
@pytest.mark.parametrize("order", [True, False])
def test_index_narrow_kernel_shapes(C, order):
    """Lane-per-row kernel (index_narrow.cu): aligned tile bases vs arbitrary run
    starts, one-row records, dense tiles (several passes), tile sizes, and the
    generic warp-stream kernel on the same input."""
    lens = [1, 5000, 2, 1, 3, 12001, 300, 7777, 1]
    recs = [(f"c{i}", n) for i, n in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=300 + i, dense=(i % 2 == 1))
                           for i, n in enumerate(lens)])
    want = mo.index_build(vals, recs, order)
    for kw in ({}, {"rows_per_tile": 1}, {"rows_per_tile": 512}, {"rows_per_tile": 900, "stages": 1, "warps_per_cta": 4},
               {"stages": 4, "warps_per_cta": 3}, {"kernel_variant": 1}, {"kernel_variant": 3},
               {"kernel_variant": 3, "rows_per_tile": 5, "emit_buf_records": 7}):
        res, got = gpu_index(vals, recs, order, **kw)
        assert_index_equal(got, want, f"narrow C={C} order={order} {kw}")
        assert not res.general


@pytest.mark.parametrize("C", [4, 9, 10, 40, 93])
@pytest.mark.parametrize("shift", [1, 2, 3, 4, 7])
def test_index_narrow_kernel_position_shards(C, shift):
    """A position shard: the buffer starts `shift` rows before the owned rows (row
    shift-1 is the halo), the run continues a record and is continued by another
    shard (no chr-end rows), followed by halo rows that belong to nobody."""
    api = _api()
    L = 9000
    vals = mo.synth_dap(L, C, seed=77 + C, dense=True)
    whole = mo.index_build(vals, [("chrS", L)], True)
    lo, hi = 2500, 7013
    buf = torch.from_numpy(np.ascontiguousarray(vals[lo - shift:hi + 50], dtype=np.int32)).cuda()
    segs = [api.Segment(row_begin=shift, n_rows=hi - lo, pos0=lo, rec_len=L, rec_id=0, flags=0)]
    res = api.IndexBuilder().build(buf, C, segs, True)
    got = res.to_host()
    keep = (whole[1] >= lo) & (whole[1] < hi)
    assert_index_equal(got, tuple(w[keep] for w in whole), f"shard shift={shift}")


def test_index_position_beyond_records_raises():
    api = _api()
    dap = torch.zeros((10, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(Exception, match="beyond all intervals"):
        api.index_build(dap, [("a", 4)], True)


def test_index_large_values_uint32_end():
    # MEM end = p + length may exceed int32; ends are carried as uint32
    C, L = 5, 300
    vals = np.full((L, C), 2**31 - 1 - 100, dtype=np.int64)
    vals[::7, 2] = 2**31 - 1
    want = mo.index_build(vals, [("big", L)], True)
    res, got = gpu_index(vals, [("big", L)], True)
    assert_index_equal(got, want, "u32")


GRCH38 = (248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
          138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
          83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415)


@pytest.mark.parametrize("C,order", [(9, True), (9, False), (93, True)])
def test_index_whole_genome_layout_scaled(C, order):
    """BASELINE configs[4]'s layout at 1/2000 scale (C = 9) / 1/20000 (C = 93): the 24 GRCh38
    primary records in one DAP, every record with its own synthetic matching statistics;
    index rows (incl. the chr-end rows of all 24 records) bit-exact against the C port of the
    reference algorithm, and one k-mer query per record against the oracle."""
    from oracle import c_oracle as co
    api = _api()
    scale = 2000 if C == 9 else 20000
    records = [(f"chr{i + 1}", n // scale) for i, n in enumerate(GRCH38)]
    parts = [mo.synth_dap(n, C, seed=100 + i) for i, (_, n) in enumerate(records)]
    vals = np.concatenate(parts)
    want = co.index_build(vals, records, order)
    res, got = gpu_index(vals, records, order)
    assert_index_equal(got, want, "whole genome layout")
    # queries: rows of one record, whole record window (past its end as well)
    for rid in (0, 7, 23):
        m = want[0] == rid
        n_rec = records[rid][1]
        for membership in ((False,) if order else (True,)):
            w = mo.query(want[1][m], want[2][m], want[3][m], 0, n_rec + 50, 31, C + 1, membership)
            g, _ = gpu_query(want[1][m], want[2][m], want[3][m], 0, n_rec + 50, 31, C + 1, membership)
            assert np.array_equal(g, w), (rid, membership)


# ------------------------------------------------------------------ queries
@pytest.mark.parametrize("C,membership", [(9, False), (9, True), (93, False), (93, True), (40, True)])
def test_query_vs_oracle(C, membership):
    api = _api()
    rng = np.random.default_rng(C + membership)
    L = 60000
    vals = mo.synth_dap(L, C, seed=C + 11, dense=True)
    recs = [("chrQ", L)]
    _, s, e, c = mo.index_build(vals, recs, not membership)
    n_docs = C + 1
    windows = [(0, L), (0, 1), (17, 8192 + 17), (L - 5, L + 40), (12345, 12345), (30000, 47001)]
    for qs, qe in windows:
        for k in (1, 2, 31, 101, 300):
            want = mo.query(s, e, c, qs, qe, k, n_docs, membership)
            got, dev = gpu_query(s, e, c, qs, qe, k, n_docs, membership)
            assert np.array_equal(got, want), (qs, qe, k)
    # text formatters, byte-identical to the reference's writers
    want = mo.query(s, e, c, 100, 9000, 31, n_docs, membership)
    got, dev = gpu_query(s, e, c, 100, 9000, 31, n_docs, membership)
    if membership:
        assert api.format_membership(dev, n_docs).decode() == mo.format_membership(want)
    else:
        assert api.format_conservation(dev).decode() == mo.format_conservation(want)


def test_query_invariant_valid_ms():
    # conservation == 1 + #{MS >= k} for valid matching statistics
    L, C, k = 300000, 9, 31
    api = _api()
    dap = api.synth_dap(L, C, seed=20240612)
    res = api.index_build(dap, [("chrS", L)], True)
    out = api.query_conservation(res.start[:res.n], res.end[:res.n], res.order[:res.n], 0, L, k, C + 1)
    want = 1 + (dap >= k).sum(dim=1)
    assert torch.equal(out.to(torch.int64), want)


@pytest.mark.parametrize("density,n_docs", [(3.0, 12), (0.3, 200), (0.002, 300), (0.05, 95)])
def test_query_arbitrary_rows(density, n_docs):
    """Index rows that no index build would produce (any parquet can be queried):
    random ends >= start, unordered orders and ends within a start position,
    long empty stretches; conservation (uint8 / uint16) and membership."""
    rng = np.random.default_rng(int(density * 1000) + n_docs)
    span = 200000
    n = max(3, int(span * density))
    s = np.sort(rng.integers(1, span, size=n))
    e = s + rng.integers(0, 80, size=n) * (rng.random(n) < 0.8) + rng.integers(0, 5000, size=n) * (rng.random(n) < 0.1)
    for membership in (False, True):
        c = rng.integers(0, n_docs if membership else n_docs + 1, size=n)
        for qs, qe in ((0, span + 3000), (span // 3, span // 3 + 70001), (5, 1029), (span - 10, span + 5000)):
            for k in (2, 31, 64):
                want = mo.query(s, e, c, qs, qe, k, n_docs, membership)
                got, _ = gpu_query(s, e, c, qs, qe, k, n_docs, membership)
                if not membership:
                    got = got.astype(np.int64) & 0xFFFF
                assert np.array_equal(got, want), (membership, qs, qe, k)


@pytest.mark.parametrize("C,membership", [(9, False), (40, False), (93, True)])
def test_query_position_shards_and_k_sweep(C, membership):
    """SURVEY 8e / BASELINE configs[4]: window positions split into ranges, every range
    answered from its own rows plus a right halo of k_max - 1 positions, concatenated in
    range order == the whole-window result, for a sweep of k."""
    from memo_b200 import host, shard
    L = 50000
    vals = mo.synth_dap(L, C, seed=3 * C + 1, dense=True)
    _, s, e, c = mo.index_build(vals, [("chrK", L)], not membership)
    n_docs = C + 1
    ks = (15, 31, 64, 101)
    whole = host.query_sweep(s, e, c, 0, L, ks, n_docs, membership)
    for k in ks:
        want = mo.query(s, e, c, 0, L, k, n_docs, membership)
        assert np.array_equal(whole[k].astype(np.int64), want), k
    for world in (2, 3, 8):
        parts = {k: [] for k in ks}
        for rank in range(world):
            lo, hi = shard.shard_range(L, world, rank)
            a, b = shard.query_rows_for_range(s, lo, hi, max(ks))
            got = host.query_sweep(s[a:b], e[a:b], c[a:b], lo, hi, ks, n_docs, membership)
            for k in ks:
                parts[k].append(got[k])
        for k in ks:
            assert np.array_equal(np.concatenate(parts[k]), whole[k]), (world, k)


@pytest.mark.parametrize("C,membership", [(9, False), (93, False), (93, True)])
def test_query_sweep_one_launch_many_k(C, membership):
    """memo_query_sweep: 20 k values (two launches of <= 16) incl. k = 1, 2 and a window whose
    start is the dense start of the record (heavy tiles handed on in pieces answer every k),
    against the oracle per k; unaligned window length."""
    api = _api()
    L = 120_000
    vals = mo.synth_dap(L, C, seed=5 * C + 2)
    _, s, e, c = mo.index_build(vals, [("chrK", L)], not membership)
    n_docs = C + 1
    f1 = torch.from_numpy(s.astype(np.int32)).cuda()
    f2 = torch.from_numpy(e.astype(np.uint32).view(np.int32)).cuda()
    f3 = torch.from_numpy(c.astype(np.int32)).cuda()
    ks = [1, 2, 3, 15, 16, 17, 21, 31, 32, 33, 41, 51, 61, 64, 65, 71, 81, 91, 101, 129]
    lo, hi = 0, (L - 1000 if membership else L - 1003)
    out = api.query_sweep(f1, f2, f3, lo, hi, ks, n_docs, membership).cpu().numpy()
    for i, k in enumerate(ks):
        want = mo.query(s, e, c, lo, hi, k, n_docs, membership)
        got = api.unpack_membership(out[i], n_docs) if membership else out[i, :hi - lo].astype(np.int64)
        assert np.array_equal(got, want), k


def test_query_order_out_of_range_raises():
    s = np.array([5, 9]); e = np.array([9, 12]); c = np.array([1, 7])
    with pytest.raises(IndexError):
        gpu_query(s, e, c, 0, 20, 3, 5, False)
    with pytest.raises(IndexError):
        gpu_query(s, e, c, 0, 20, 3, 5, True)


def test_query_u16_many_docs():
    s = np.array([5, 9, 14]); e = np.array([5, 9, 15]); c = np.array([300, 2, 999])
    want = mo.query(s, e, c, 0, 30, 4, 1000, False)
    got, _ = gpu_query(s, e, c, 0, 30, 4, 1000, False)
    assert np.array_equal(got.astype(np.int64) & 0xFFFF, want)


# ------------------------------------------------------------------ host-buffer API and CLIs
@pytest.mark.parametrize("order", [True, False])
def test_host_build_index_streams_in_chunks(order):
    from memo_b200 import host
    C = 9
    lens = [30000, 1, 7, 52000, 300]
    recs = [(f"c{i}", n) for i, n in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=7 + i, dense=(i == 3)) for i, n in enumerate(lens)])
    want = mo.index_build(vals, recs, order)
    for chunk_bytes in (64 << 20, 100_000, 36 * 1000):          # 1, ~30 and ~82 chunks; dense run overflows
        stats = {}
        got = host.build_index(vals.astype(np.int32), recs, order, chunk_bytes=chunk_bytes, stats=stats)
        assert_index_equal(got, want, f"chunk_bytes={chunk_bytes}")
        assert not stats["general"]
    raw = host.build_index(vals.astype(np.int32), recs, order, chunk_bytes=100_000, raw=True)
    assert raw.start.dtype == np.int32 and raw.end.dtype == np.uint32 and raw.n == want[1].size
    assert_index_equal(raw.as_int64(), want, "raw")


def test_host_build_index_irregular_falls_back():
    from memo_b200 import host
    rng = np.random.default_rng(3)
    vals = rng.integers(0, 40, (5000, 6))
    recs = [("a", 3000), ("b", 2000)]
    stats = {}
    got = host.build_index(vals.astype(np.int32), recs, True, chunk_bytes=20_000, stats=stats)
    assert stats["general"]
    assert_index_equal(got, mo.index_build(vals, recs, True), "irregular")


def _stream_rows(blocks, recs, order, **kw):
    from memo_b200 import host
    parts, stats = [], {}

    def on_rows(rec_counts, start, end, col):
        rec = np.repeat([r for r, _ in rec_counts], [c for _, c in rec_counts]).astype(np.int64)
        parts.append((rec, start.astype(np.int64), end.astype(np.int64), col.astype(np.int64)))

    n = host.build_index_streaming(blocks, recs, order, on_rows, stats=stats, **kw)
    got = tuple(np.concatenate([p[i] for p in parts]) if parts else np.zeros(0, dtype=np.int64) for i in range(4))
    assert n == got[1].size
    return got, stats


@pytest.mark.parametrize("order", [True, False])
def test_index_stream_blocks_of_any_size(order):
    """The streaming pipeline behind the dap_to_bed drop-in: blocks of arbitrary sizes in, index
    rows out chunk by chunk; chunk cuts inside records, at record ends, one-row records, a DAP
    that stops inside its last record, pinned blocks taken in place."""
    C = 7
    lens = [3000, 1, 4500, 1, 1, 2500]
    recs = [(f"c{i}", n) for i, n in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=50 + i, dense=(i % 2 == 0)) for i, n in enumerate(lens)])
    vals = np.ascontiguousarray(vals[:-37], dtype=np.int32)
    want = mo.index_build(vals, recs, order)
    rng = np.random.default_rng(4)
    for chunk_rows in (1, 333, 3000, 3001, 5000, 10 ** 6):
        cuts = np.sort(rng.integers(0, len(vals), 9))
        blocks = [b for b in np.split(vals, cuts)]
        got, stats = _stream_rows(iter(blocks), recs, order, chunk_bytes=4 * C * chunk_rows)
        assert_index_equal(got, want, f"chunk_rows={chunk_rows}")
        assert not stats["general"]
    pinned = torch.from_numpy(vals).pin_memory()
    got, _ = _stream_rows(iter([pinned[:4000], pinned[4000:4001], pinned[4001:]]), recs, order, chunk_bytes=4 * C * 1500)
    assert_index_equal(got, want, "pinned blocks")
    # a stream that starts inside a record (dap.txt whose first position is not 0)
    got, _ = _stream_rows(iter([vals[700:]]), recs, order, chunk_bytes=4 * C * 1000, pos_first=700)
    assert_index_equal(got, mo.index_build(vals[700:], recs, order, pos=np.arange(700, len(vals))), "pos_first")


@pytest.mark.parametrize("order", [True, False])
def test_index_stream_turns_exact_where_input_turns_irregular(order):
    """Valid matching statistics, then arbitrary integers from the middle of a record on: the
    chunks before stay single-pass, the rest runs the exact build with the carry handed on."""
    C = 5
    rng = np.random.default_rng(8)
    valid = mo.synth_dap(9000, C, seed=3, dense=True)
    junk = rng.integers(0, 60, (6000, C))
    junk[rng.random(junk.shape) < 0.4] = 0
    junk[:, 0] = np.maximum(7000 - 2 * np.arange(6000), 0)          # carries that cross many chunks
    vals = np.ascontiguousarray(np.concatenate([valid[:5000], junk, valid[5000:]]), dtype=np.int32)
    recs = [("a", 8000), ("b", 4000), ("c", 3000)]
    want = mo.index_build(vals, recs, order)
    for chunk_rows in (700, 2048, 10 ** 6):
        got, stats = _stream_rows(iter([vals]), recs, order, chunk_bytes=4 * C * chunk_rows)
        assert_index_equal(got, want, f"chunk_rows={chunk_rows}")
        assert stats["general"] and (chunk_rows > 5000 or stats["chunks_general"] < stats["chunks"])


def test_index_stream_memory_is_bounded_by_the_chunk():
    """3 M rows through 1 MB chunks: device memory stays O(chunk), not O(DAP)."""
    from memo_b200 import api
    L, C = 3_000_000, 9
    dap = api.synth_dap(L, C, seed=11).cpu().numpy()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    got, stats = _stream_rows((dap[a:a + 250_000] for a in range(0, L, 250_000)), [("chrS", L)], True,
                              chunk_bytes=1 << 20)
    peak = torch.cuda.max_memory_allocated() - base
    assert stats["chunks"] >= 100 and peak < 32 << 20, (stats, peak)
    want_head = mo.index_build(dap[:200_000], [("chrS", 200_000)], True)
    m = got[1] < 199_000
    assert np.array_equal(got[1][m], want_head[1][want_head[1] < 199_000])


def test_dap_text_parsed_on_the_device(tmp_path):
    """memo_dap_text_parse (dap.txt -> int32 rows on the GPU) against numpy: blocks cut at line
    ends, a last line without newline, byte shares of a file, and everything int() would reject."""
    from memo_b200 import io
    from memo_b200._lib import MemoError
    rng = np.random.default_rng(12)
    L, C = 30_011, 7
    vals = rng.integers(0, 3_000_000, (L, C))
    vals[rng.random((L, C)) < 0.3] = rng.integers(0, 10, int((rng.random((L, C)) < 0.3).sum() * 0 + 1))[0]
    vals[5, 2] = 2**31 - 1
    text = "".join(f"{i + 40} " + " ".join(map(str, row)) + "\n" for i, row in enumerate(vals))
    path = tmp_path / "dap.txt"
    path.write_text(text)

    def collect(p, **kw):
        parts = [(p0, b.cpu().numpy()) for p0, b in io.iter_dap_text_device(str(p), **kw)]
        return parts

    for block in (1 << 12, 77_777, 1 << 26):
        parts = collect(path, block_bytes=block)
        assert parts[0][0] == 40 and np.array_equal(np.concatenate([b for _, b in parts]), vals), block
        pos = 40
        for p0, b in parts:
            assert p0 == pos
            pos += len(b)
    nonl = tmp_path / "nonl.txt"
    nonl.write_text(text[:-1])
    assert np.array_equal(np.concatenate([b for _, b in collect(nonl, block_bytes=50_000)]), vals)
    rows, expect = [], 40
    for r in range(3):
        lo, hi, prev = io.split_text_rows(str(path), 3, r)
        for p0, b in collect(path, block_bytes=20_000, byte_range=(lo, hi)):
            assert p0 == expect
            expect += len(b)
            rows.append(b)
    assert np.array_equal(np.concatenate(rows), vals)
    bad = tmp_path / "bad.txt"
    for body, exc in (("0 1 2\n1 3 x\n", ValueError), ("0 1 2\n1  3\n", ValueError), ("0 1 2\n\n1 3 4\n", ValueError),
                      ("0 1 2\n1 3\n", MemoError), ("0 1 2\n1 3 4 5\n", MemoError), ("0 1 2\n2 3 4\n", MemoError),
                      ("0 1 2\n1 3 2147483648\n", MemoError), ("0 1 -2\n", ValueError)):
        bad.write_text(body)
        with pytest.raises(exc):
            collect(bad)
    bad.write_text("")
    assert collect(bad) == []


def test_view_bins_match_oracle(tmp_path):
    """`memo view` binning kernel (plot_conservation.py:46-58) against the oracle, through the
    API and through the drop-in's preprocess_data; uint8 and uint16 vectors, bins that do not
    divide the window, more bins than a slice, values above -n."""
    from memo_b200 import api, plot_conservation
    rng = np.random.default_rng(6)
    for n, n_docs, n_bins in ((20, 5, 4), (100_003, 10, 500), (300_000, 94, 7), (70_001, 400, 33), (5000, 3, 5000)):
        vec = rng.integers(1, n_docs + 1, n)
        vec[rng.random(n) < 0.7] = n_docs
        if n == 20:
            vec[3] = n_docs + 2                                # not in the table, still in the bin size
        want = mo.view_bins(vec, n_docs, n_bins)
        wide = n_docs > 255
        dev = torch.from_numpy(vec.astype(np.int16 if wide else np.uint8)).cuda()
        got = plot_conservation.bin_composition(dev, n_docs, n_bins)
        assert np.array_equal(got, want), (n, n_docs, n_bins)
        counts = api.view_bins(dev, n_docs, n_bins)
        assert int(counts.sum()) == int((vec <= n_docs).sum())
    path = tmp_path / "cons.txt"
    path.write_text("\n".join(map(str, vec)) + "\n")
    df = plot_conservation.preprocess_data(str(path), n_docs, n_bins)
    assert df["value"].tolist() == want[:, :n_docs].T.reshape(-1).tolist()
    assert df["bin"].tolist() == np.tile(np.arange(n_bins), n_docs).tolist()
    with pytest.raises(ZeroDivisionError):
        plot_conservation.bin_composition(dev[:3], n_docs, 10)


def test_host_query_matches_oracle():
    from memo_b200 import host
    C, L = 9, 40000
    vals = mo.synth_dap(L, C, seed=5, dense=True)
    _, s, e, c = mo.index_build(vals, [("q", L)], True)
    want = mo.query(s, e, c, 10, L - 3, 31, C + 1, False)
    assert np.array_equal(host.query(s, e, c, 10, L - 3, 31, C + 1, False), want)
    raw = host.query(s.astype(np.int32), e.astype(np.uint32), c.astype(np.int32), 10, L - 3, 31, C + 1,
                     False, raw=True, trusted=True)
    assert raw.dtype == np.uint8 and np.array_equal(raw.astype(np.int64), want)
    # unsorted rows are accepted (painting is order independent)
    perm = np.random.default_rng(0).permutation(s.size)
    assert np.array_equal(host.query(s[perm], e[perm], c[perm], 10, L - 3, 31, C + 1, False), want)


def test_cli_round_trip_example(example_golden, tmp_path, capsysbinary):
    """dap.txt -> BED (dap_to_bed) -> Parquet (parquet_compress_bed) -> query text
    (memo_query), byte-identical to the reference-generated goldens."""
    from memo_b200 import dap_to_bed, memo_query, parquet_compress_bed
    g = example_golden
    dap = tmp_path / "dap.txt"; dap.write_text(g["dap_txt"])
    fai = tmp_path / "ref_1.fa.fai"
    fai.write_text("".join(f"{h}\t{n}\t7\t{n}\t{n + 1}\n" for h, n in g["records"]))
    for order, bed_key, tag in ((True, "cons_bed", "cons"), (False, "memb_bed", "memb")):
        argv = ["--mem", "--overlap", "--fai", str(fai), "--dap", str(dap)] + (["--order"] if order else [])
        args = dap_to_bed.parse_arguments(argv)
        dap_to_bed.check_args(args)
        bed = tmp_path / f"{tag}.bed"
        with open(bed, "wb") as fh:
            dap_to_bed.main(args, sink=fh)
        assert bed.read_text() == g[bed_key]
        # extension: the per-genome MONI .lengths files instead of dap.txt
        from test_host_logic import _write_lengths
        lens = _write_lengths(tmp_path, g["vals"])
        args = dap_to_bed.parse_arguments(["--mem", "--overlap", "--fai", str(fai), "--lengths"] + lens +
                                          (["--order"] if order else []))
        dap_to_bed.check_args(args)
        with open(tmp_path / f"{tag}_lengths.bed", "wb") as fh:
            dap_to_bed.main(args, sink=fh)
        assert (tmp_path / f"{tag}_lengths.bed").read_text() == g[bed_key]
        pq_path = tmp_path / f"{tag}.parquet"
        parquet_compress_bed.main(parquet_compress_bed.parse_arguments(["-f", str(bed), "-o", str(pq_path)]))
        capsysbinary.readouterr()
        # extension: the Parquet index straight from the device columns (no BED text in between)
        import pyarrow.parquet as pq
        direct = tmp_path / f"{tag}_direct.parquet"
        args = dap_to_bed.parse_arguments(argv + ["--parquet", str(direct)])
        dap_to_bed.check_args(args)
        dap_to_bed.main(args)
        assert pq.read_table(direct).equals(pq.read_table(pq_path))
        for q in g["queries"]:
            if q["membership"] != (not order):
                continue
            out = tmp_path / "q.txt"
            argv = ["-b", str(pq_path), "-r", q["region"], "-k", str(q["k"]), "-n", str(q["n"]), "-o", str(out)]
            memo_query.main(memo_query.parse_arguments(argv + (["-m"] if q["membership"] else [])))
            assert out.read_text() == q["out"], q


def test_bed_formatter_matches_python_text():
    """memo_format_bed against '\\t'.join(map(str, ...)) (src/dap_to_bed.py:105): every digit count
    of start / end / order, end beyond int32, empty and long record names, ragged block tails."""
    import torch
    from memo_b200 import api
    rng = np.random.default_rng(77)
    fmt = api.BedFormatter()
    for n, name in ((0, "chr1"), (1, ""), (7, "c"), (2049, "chr1"), (50_000, "HG002#1#JAHKSE010000012.1" * 8)):
        mag = rng.integers(0, 10, n)
        start = np.minimum((10.0 ** mag * rng.random(n)).astype(np.int64), 2**31 - 1)
        end = np.minimum(start + (10.0 ** rng.integers(0, 11, n) * rng.random(n)).astype(np.int64), 2**32 - 1)
        order = np.minimum((10.0 ** rng.integers(0, 4, n) * rng.random(n)).astype(np.int64) + 1, 512)
        if n > 3:
            start[:3] = (0, 2**31 - 1, 999_999_999)
            end[:3] = (0, 2**32 - 1, 1_000_000_000)
            order[:3] = (1, 512, 10)
        rows = torch.from_numpy(np.stack([start.astype(np.int32), end.astype(np.uint32).view(np.int32),
                                          order.astype(np.int32)])).cuda()
        want = "".join("\t".join(map(str, [name, s, e, o])) + "\n" for s, e, o in zip(start, end, order))
        assert bytes(fmt.format(rows, name)).decode("utf-8") == want


def test_cli_bed_text_device_equals_host_formatter(tmp_path, monkeypatch):
    """dap_to_bed writes the same bytes with the device BED formatter and with the Arrow CSV
    writer (MEMO_BED_FORMAT=host), several records and chunks."""
    from memo_b200 import dap_to_bed
    rng = np.random.default_rng(5)
    recs = [("chrA", 7000), ("chrB_random", 3000), ("c", 1)]
    L, C = sum(n for _, n in recs), 5
    vals = mo.synth_dap(7000, C, 3)
    vals = np.concatenate([vals, mo.synth_dap(3000, C, 4), mo.synth_dap(1, C, 5)])
    dap = tmp_path / "dap.txt"
    dap.write_text("".join(" ".join(map(str, [i] + list(r))) + "\n" for i, r in enumerate(vals)))
    fai = tmp_path / "x.fa.fai"
    fai.write_text("".join(f"{h}\t{n}\t7\t{n}\t{n + 1}\n" for h, n in recs))
    monkeypatch.setenv("MEMO_CHUNK_BYTES", str(40_000))          # many chunks
    outs = {}
    for mode in ("device", "host"):
        monkeypatch.setenv("MEMO_BED_FORMAT", mode)
        args = dap_to_bed.parse_arguments(["--mem", "--overlap", "--order", "--fai", str(fai), "--dap", str(dap)])
        dap_to_bed.check_args(args)
        with open(tmp_path / f"{mode}.bed", "wb") as fh:
            dap_to_bed.main(args, sink=fh)
        outs[mode] = (tmp_path / f"{mode}.bed").read_bytes()
    assert outs["device"] == outs["host"] and len(outs["device"]) > 0
    assert outs["device"].decode() == mo.format_bed(recs, *mo.index_build(vals, recs, True))


@pytest.mark.parametrize("C", [9, 93])
def test_index_more_record_runs_than_one_prep_launch(C):
    """150 short records in one buffer: the record runs reach the device as kernel parameters,
    64 per prep launch (index_build.cu) -- three launches here."""
    lens = [37 + (i * 13) % 41 for i in range(150)]
    recs = [(f"r{i}", n) for i, n in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=900 + i, dense=(i % 3 == 0)) for i, n in enumerate(lens)])
    for order in (True, False):
        res, got = gpu_index(vals, recs, order)
        assert not res.general
        assert_index_equal(got, mo.index_build(vals, recs, order), f"C={C} order={order} 150 records")
