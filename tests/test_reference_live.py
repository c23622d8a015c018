"""Live differential: the oracle against the UNMODIFIED reference scripts, run as child
processes where the reference tree exists (the build container: /root/reference, or
$MEMO_REFERENCE).  Skipped elsewhere -- the committed goldens (tests/golden/) carry the
same comparison to the GPU box.  Fresh seeds every time the file changes, so this is a
second, independent pin of the oracle next to the goldens."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import memo_oracle as mo

REF = os.environ.get("MEMO_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(SRC, "dap_to_bed.py")),
                                reason="reference tree not present")


def _random_case(rng, valid):
    n_rec = int(rng.integers(1, 4))
    lens = [int(rng.integers(2, 40)) for _ in range(n_rec)]
    C = int(rng.integers(1, 7))
    L = sum(lens)
    if valid:
        parts = [mo.synth_dap(n, C, seed=int(rng.integers(1, 1 << 30)), dense=True) for n in lens]
        vals = np.concatenate(parts)
    else:
        vals = rng.integers(0, 12, size=(L, C))
    return [(f"rec{i}", n) for i, n in enumerate(lens)], vals.astype(np.int64)


def _run_reference_index(tmp_path, records, vals, order):
    fai = tmp_path / "pivot.fa.fai"
    fai.write_text("".join(f"{h}\t{n}\t7\t{n}\t{n + 1}\n" for h, n in records))
    dap = tmp_path / "dap.txt"
    dap.write_text("".join(f"{p} " + " ".join(str(v) for v in row) + "\n" for p, row in enumerate(vals)))
    argv = [sys.executable, os.path.join(SRC, "dap_to_bed.py"), "--mem", "--overlap", "--fai", str(fai),
            "--dap", str(dap)] + (["--order"] if order else [])
    return subprocess.run(argv, check=True, capture_output=True, text=True).stdout


def _bed_text(records, rows):
    rec, s, e, c = rows
    return "".join(f"{records[r][0]}\t{a}\t{b}\t{o}\n" for r, a, b, o in zip(rec, s, e, c))


@pytest.mark.parametrize("seed", range(8))
def test_index_against_live_reference(seed, tmp_path):
    rng = np.random.default_rng(7000 + seed)
    records, vals = _random_case(rng, valid=seed % 2 == 0)
    for order in (True, False):
        want = _run_reference_index(tmp_path, records, vals, order)
        assert _bed_text(records, mo.index_build(vals, records, order)) == want, (seed, order)


def test_query_against_live_reference(tmp_path):
    pytest.importorskip("numba")
    pa = pytest.importorskip("pyarrow")
    import pyarrow.parquet as pq
    rng = np.random.default_rng(99)
    L, C = 400, 5
    records = [("chrQ", L)]
    vals = mo.synth_dap(L, C, seed=4242, dense=True).astype(np.int64)
    for membership in (False, True):
        rec, s, e, c = mo.index_build(vals, records, not membership)
        table = pa.table({"f0": pa.array(["chrQ"] * len(s), pa.utf8()), "f1": pa.array(s, pa.int64()),
                          "f2": pa.array(e, pa.int64()), "f3": pa.array(c, pa.int64())})
        pq_path = tmp_path / ("m.parquet" if membership else "c.parquet")
        pq.write_table(table, pq_path, compression="ZSTD")
        for (qs, qe, k) in ((0, L, 31), (37, 311, 5), (L - 9, L + 20, 12)):
            out = tmp_path / "q.txt"
            argv = [sys.executable, os.path.join(SRC, "memo_query.py"), "-b", str(pq_path), "-r", f"chrQ:{qs}-{qe}",
                    "-k", str(k), "-n", str(C + 1), "-o", str(out)] + (["-m"] if membership else [])
            subprocess.run(argv, check=True, capture_output=True, text=True)
            got = mo.query(s, e, c, qs, qe, k, C + 1, membership)
            text = mo.format_membership(got) if membership else mo.format_conservation(got)
            assert out.read_text() == text, (membership, qs, qe, k)


def test_parquet_stage_against_live_reference(tmp_path):
    """parquet_compress_bed.py of the reference and of memo_b200 on the same BED: equal
    tables, equal schema (names, types, no schema metadata), ZSTD (table
    equality is the contract, the file embeds a writer version string)."""
    pytest.importorskip("pandas")
    pq = pytest.importorskip("pyarrow.parquet")
    rng = np.random.default_rng(5)
    records, vals = _random_case(rng, valid=True)
    bed = tmp_path / "idx.bed"
    bed.write_text(_bed_text(records, mo.index_build(vals, records, True)))
    ref_out, our_out = tmp_path / "ref.parquet", tmp_path / "our.parquet"
    subprocess.run([sys.executable, os.path.join(SRC, "parquet_compress_bed.py"), "-f", str(bed), "-o", str(ref_out)],
                   check=True, capture_output=True, text=True)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-m", "memo_b200.parquet_compress_bed", "-f", str(bed), "-o", str(our_out)],
                   check=True, capture_output=True, text=True, cwd=root)
    a, b = pq.read_table(str(ref_out)), pq.read_table(str(our_out))
    assert a.schema.equals(b.schema, check_metadata=True), (a.schema, b.schema)
    assert a.equals(b)
    ma, mb = pq.ParquetFile(str(ref_out)).metadata, pq.ParquetFile(str(our_out)).metadata
    assert ma.row_group(0).column(1).compression == mb.row_group(0).column(1).compression == "ZSTD"


@pytest.mark.parametrize("seed", range(24))
def test_index_edge_shapes_against_live_reference(seed, tmp_path):
    """Wider shapes than the cases above: records of one row, one column, DAPs that stop inside a
    record (the chr-end rows still follow the last row, dap_to_bed.py:133-134), values far beyond
    the record length, rows of zeros and ties -- oracle (numpy) and C port against the script."""
    from oracle import c_oracle as co
    rng = np.random.default_rng(8100 + seed)
    lens = [int(rng.integers(1, 30)) for _ in range(int(rng.integers(1, 5)))]
    C, L = int(rng.integers(1, 8)), sum(lens)
    mode = seed % 4
    if mode == 0:
        vals = np.concatenate([mo.synth_dap(n, C, seed=int(rng.integers(1, 1 << 30)), dense=True) for n in lens])
    else:
        vals = rng.integers(0, (12, 3, 80)[mode - 1], size=(L, C))
    if rng.random() < 0.4 and L > 1:
        vals = vals[:int(rng.integers(1, L))]                     # partial DAP
    vals = vals.astype(np.int64)
    records = [(f"r{i}", n) for i, n in enumerate(lens)]
    for order in (True, False):
        want = _run_reference_index(tmp_path, records, vals, order)
        assert _bed_text(records, mo.index_build(vals, records, order)) == want, (seed, order)
        assert _bed_text(records, co.index_build(vals, records, order)) == want, (seed, order, "C port")


@pytest.mark.parametrize("seed", range(6))
def test_query_edge_windows_against_live_reference(seed, tmp_path):
    """memo_query.py on a multi-record Parquet index: k from 1 to beyond the window, windows that
    start anywhere and end beyond the record, indexes of regular and irregular DAPs."""
    pytest.importorskip("numba")
    pa = pytest.importorskip("pyarrow")
    import pyarrow.parquet as pq
    rng = np.random.default_rng(9100 + seed)
    lens = [int(rng.integers(20, 300)) for _ in range(int(rng.integers(1, 4)))]
    C, L = int(rng.integers(1, 8)), sum(lens)
    if seed % 3 == 0:
        vals = np.concatenate([mo.synth_dap(n, C, seed=int(rng.integers(1, 1 << 30)), dense=True) for n in lens])
    else:
        vals = rng.integers(0, (40, 5)[seed % 3 - 1], size=(L, C))
    records = [(f"r{i}", n) for i, n in enumerate(lens)]
    membership = bool(seed & 1)
    rec, s, e, c = mo.index_build(vals.astype(np.int64), records, not membership)
    table = pa.table({"f0": pa.array([records[r][0] for r in rec], pa.utf8()), "f1": pa.array(s, pa.int64()),
                      "f2": pa.array(e, pa.int64()), "f3": pa.array(c, pa.int64())})
    pq_path = tmp_path / "idx.parquet"
    pq.write_table(table, pq_path, compression="ZSTD")
    ri = int(rng.integers(0, len(lens)))
    qs = int(rng.integers(0, lens[ri]))
    qe = int(rng.integers(qs + 1, lens[ri] + 40))
    k = (1, 2, 5, 31, 64, 200)[seed]
    out = tmp_path / "q.txt"
    argv = [sys.executable, os.path.join(SRC, "memo_query.py"), "-b", str(pq_path), "-r", f"r{ri}:{qs}-{qe}",
            "-k", str(k), "-n", str(C + 1), "-o", str(out)] + (["-m"] if membership else [])
    subprocess.run(argv, check=True, capture_output=True, text=True)
    m = rec == ri
    got = mo.query(s[m], e[m], c[m], qs, qe, k, C + 1, membership)
    text = mo.format_membership(got) if membership else mo.format_conservation(got)
    assert out.read_text() == text, (membership, ri, qs, qe, k)


@pytest.mark.parametrize("C,order", [(9, True), (9, False), (40, True)])
def test_c_port_against_live_reference_at_scale(C, order, tmp_path):
    """The C port (bench.py's cpu_baseline / --impl reference arm) on a synthetic
    HPRC-shaped DAP of 40 k rows in two records, against the reference script itself."""
    from oracle import c_oracle as co
    lens = [25000, 15000]
    records = [(f"chr{i + 1}", n) for i, n in enumerate(lens)]
    vals = np.concatenate([mo.synth_dap(n, C, seed=31 + i) for i, n in enumerate(lens)]).astype(np.int64)
    want = _run_reference_index(tmp_path, records, vals, order)
    assert _bed_text(records, co.index_build(vals, records, order)) == want


def _reference_preprocess_data():
    """src/plot_conservation.py's preprocess_data, imported from the unmodified file with a stub
    standing in for plotnine (absent here; the function itself only needs numpy / pandas)."""
    import importlib.util
    import types
    stub = types.ModuleType("plotnine")
    for name in ("ggplot aes theme themes element_blank element_line element_text geom_bar ggtitle xlab ylab "
                 "scale_y_continuous scale_fill_gradient").split():
        setattr(stub, name, object())
    opts = types.ModuleType("plotnine.options")
    opts.figure_size = None
    saved = {k: sys.modules.get(k) for k in ("plotnine", "plotnine.options")}
    sys.modules["plotnine"], sys.modules["plotnine.options"] = stub, opts
    try:
        spec = importlib.util.spec_from_file_location("ref_plot_conservation", os.path.join(SRC, "plot_conservation.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.preprocess_data


def test_oracle_view_bins_vs_live_reference(tmp_path):
    """`memo view` binning: the oracle's view_bins against preprocess_data of the reference."""
    ref = _reference_preprocess_data()
    rng = np.random.default_rng(20240622)
    for case in range(12):
        n_docs = int(rng.integers(2, 12))
        n = int(rng.integers(30, 400))
        n_bins = int(rng.integers(1, 25))
        vec = rng.integers(1, n_docs + 1, n)
        vec[rng.random(n) < 0.6] = n_docs
        path = tmp_path / f"cons{case}.txt"
        path.write_text("\n".join(map(str, vec)) + "\n")
        df = ref(str(path), n_docs, n_bins)
        comp = mo.view_bins(vec, n_docs, n_bins)
        assert df["bin"].tolist() == np.tile(np.arange(n_bins), n_docs).tolist()
        assert df["No. Genomes"].tolist() == np.repeat(np.arange(n_docs, dtype=float), n_bins).tolist()
        assert df["value"].tolist() == comp[:, :n_docs].T.reshape(-1).tolist(), case
