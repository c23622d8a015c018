"""Host-side logic and the C-ABI surface (no GPU needed)."""
import io as _io
import os
import re
import subprocess
import sys

import numpy as np
import pyarrow.parquet as pq
import pytest

from oracle import memo_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from memo_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "memo_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(memo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in memo_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.memo_abi_version() == 8


def test_header_binds_from_plain_c(tmp_path):
    """The boundary is a C ABI: a C99 translation unit (no C++, no torch, no Python) includes
    include/memo_b200.h, links libmemo_b200.so and calls it.  The struct sizes the Python side
    pins are checked from the C side too."""
    import shutil
    from memo_b200 import _build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    libdir = os.path.dirname(_build.build())
    src = tmp_path / "c_abi.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "memo_b200.h"\n'
                   "int main(void) {\n"
                   "    if (sizeof(memo_segment_t) != 32 || sizeof(memo_index_opts_t) != 32) return 2;\n"
                   "    if (memo_last_error() == NULL) return 3;\n"
                   '    printf("%d\\n", memo_abi_version());\n'
                   "    return 0;\n}\n")
    exe = tmp_path / "c_abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lmemo_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.strip() == "8"


def test_struct_layouts_match_header():
    import ctypes as C
    from memo_b200 import _lib
    assert C.sizeof(_lib.Segment) == 32 and _lib.Segment.pos0.offset == 16
    assert C.sizeof(_lib.IndexOpts) == 32


def test_segments_match_oracle_runs():
    from memo_b200 import api
    rng = np.random.default_rng(3)
    for _ in range(200):
        lens = [int(x) for x in rng.integers(0, 9, int(rng.integers(1, 6)))]
        total = sum(lens)
        if total == 0:
            continue
        a = int(rng.integers(0, total))
        n = int(rng.integers(1, total - a + 1))
        recs = [(f"r{i}", l) for i, l in enumerate(lens)]
        segs = api.segments_for_rows(recs, a, n)
        runs, rel = mo.split_runs(np.arange(a, a + n), recs)
        assert [(s.rec_id, s.row_begin, s.row_begin + s.n_rows) for s in segs] == runs
        for s in segs:
            assert s.pos0 == rel[s.row_begin] and s.rec_len == lens[s.rec_id]


def test_segments_beyond_records_raise():
    from memo_b200 import api
    with pytest.raises(Exception, match="beyond all intervals"):
        api.segments_for_rows([("a", 5)], 0, 6)


def test_dap_text_ingest_and_bed_writer(example_golden, tmp_path):
    from memo_b200 import io
    p = tmp_path / "dap.txt"
    p.write_text(example_golden["dap_txt"])
    pos0, dap = io.read_dap_text(str(p))
    assert pos0 == 0 and np.array_equal(dap, example_golden["vals"])
    rows = mo.index_build(dap, example_golden["records"], True)
    buf = _io.BytesIO()
    io.write_bed(io.index_table(example_golden["records"], *rows), buf)
    assert buf.getvalue().decode() == example_golden["cons_bed"]
    bad = tmp_path / "bad.txt"
    bad.write_text("0 1 2\n2 1 2\n")
    with pytest.raises(Exception, match="consecutive"):
        io.read_dap_text(str(bad))


def _write_lengths(tmp_path, vals, per_line=(7, 1, 3)):
    """MONI-style .lengths files for the columns of `vals`: a '>' header per record and
    whitespace-separated lengths, several per line (what index.sh:79 flattens)."""
    paths = []
    for j in range(vals.shape[1]):
        n = per_line[j % len(per_line)]
        col = vals[:, j].tolist()
        lines = [">pivot_record"]
        for i in range(0, len(col), n):
            lines.append(" ".join(str(v) for v in col[i:i + n]))
            if i == 2 * n:
                lines.append(">another_record")
        p = tmp_path / f"g{j}.w_rc.lengths"
        p.write_text("\n".join(lines) + "\n")
        paths.append(str(p))
    return paths


def test_lengths_ingest_equals_dap_text(example_golden, tmp_path):
    # the per-genome MONI outputs read directly == paste | nl | re-parse
    from memo_b200 import io
    vals = example_golden["vals"]
    paths = _write_lengths(tmp_path, vals)
    assert np.array_equal(io.read_lengths_columns(paths), vals)
    vert = tmp_path / "g0.vert"
    vert.write_text("\n".join(str(v) for v in vals[:, 0]) + "\n")       # index.sh:79's vertical file
    assert np.array_equal(io.read_lengths_columns([str(vert)])[:, 0], vals[:, 0])
    short = tmp_path / "short.lengths"
    short.write_text("1 2 3\n")
    with pytest.raises(Exception, match="expected"):
        io.read_lengths_columns([paths[0], str(short)])
    bad = tmp_path / "bad.lengths"
    bad.write_text("1 x 3\n")
    with pytest.raises(ValueError):
        io.read_lengths_columns([str(bad)])


def _lengths_reference(text):
    """index.sh:79 (`grep -v '^>' | tr ' ' '\\n' | grep .`) + int() of dap_to_bed.py:87 on one file."""
    vals = []
    for line in text.split(b"\n"):
        if line.startswith(b">"):
            continue
        for tok in line.split():
            if not tok.isdigit():
                return vals, 1
            if int(tok) >= 2 ** 31:
                return vals, 8
            vals.append(int(tok))
    return vals, 0


def test_lengths_tokenizer_fuzz():
    """memo_lengths_text_parse (host code of the library) fed random runs of random text, random
    value budgets and strides, against the shell pipeline's semantics: numbers of 1 .. 10 digits
    and leading zeros, every kind of white space, headers (with digits, spaces and '>' inside)
    anywhere, bytes int() rejects, values >= 2^31, numbers cut by the end of a run."""
    import ctypes as C
    import random
    from memo_b200 import _lib
    lib = _lib.load()
    rng = random.Random(20241018)

    def run(text, stride):
        n, vals, pos, hi = len(text), [], 0, 0
        st, res = C.c_int32(0), (C.c_int64 * 3)()
        buf = (C.c_uint8 * (n + 1)).from_buffer_copy(text + b"\0")
        while True:
            if hi < n:
                hi = min(n, hi + rng.choice([1, 3, 17, 64, 100, 257, 5000]))
            final = int(hi == n)
            want = rng.choice([0, 1, 2, 5, 50, 1000, 3000])
            out = np.full((max(want, 1), stride), -7, dtype=np.int32)
            assert lib.memo_lengths_text_parse(C.addressof(buf) + pos, hi - pos, final, st, out.ctypes.data,
                                               stride, want, res) == 0
            k, err, used = res
            assert 0 <= used <= hi - pos and k <= want
            vals.extend(out[:k, 0].tolist())
            assert (out[:, 1:] == -7).all() and (out[k:] == -7).all()        # nothing but the column is written
            pos += used
            if err:
                return vals, err
            if final and pos == n:
                return vals, 0

    def text(bad):
        parts = []
        for _ in range(rng.randint(0, 300)):
            r = rng.random()
            if r < 0.03:
                parts.append(b"\n>" + bytes(rng.choice(b"abc 123>\t") for _ in range(rng.randint(0, 90))) + b"\n")
            elif r < 0.05 and bad:
                parts.append(rng.choice([b"x", b"-", b"+", b"1.5", b" >", b"\x00", b"\xff", b":", b"/"]))
            elif r < 0.06 and bad:
                parts.append(str(rng.choice([2 ** 31, 2 ** 31 + 5, 10 ** 12, 10 ** 25, 10 ** 70])).encode())
            else:
                v = min(rng.randint(0, 10 ** rng.choice([1, 1, 2, 2, 2, 3, 4, 5, 7, 8, 9, 10]) - 1), 2 ** 31 - 1)
                parts.append((b"0" * rng.randint(1, 12) if rng.random() < 0.05 else b"") + str(v).encode())
            parts.append(rng.choice([b" ", b" ", b" ", b"\n", b"\n", b"  ", b"\t", b"\r\n", b" \n ", b"\n\n", b" " * 70]))
        t = b"".join(parts)
        if rng.random() < 0.3:
            t = t.rstrip()
        return (b">hdr 1 2 3\n" if rng.random() < 0.2 else b"") + t

    n_err = 0
    for it in range(1500):
        t = text(bad=it % 4 == 0)
        want, werr = _lengths_reference(t)
        got, gerr = run(t, rng.choice([1, 1, 3, 93]))
        if werr or gerr:
            n_err += 1
            assert bool(werr) == bool(gerr), (t, werr, gerr)
        else:
            assert got == want, t
    assert n_err > 20


def test_lengths_tokenizer_every_alignment():
    """Numbers of 1 .. 10 digits at every offset of the tokenizer's 64-byte steps, followed by each
    kind of separator; a number touching the end of a run is left for the next run unless the
    file ends; a value budget stops in the middle of a step."""
    import ctypes as C
    from memo_b200 import _lib
    lib = _lib.load()

    def parse(text, final=1, want=10 ** 4):
        out, st, res = np.full(want, -7, dtype=np.int32), C.c_int32(0), (C.c_int64 * 3)()
        buf = (C.c_uint8 * (len(text) + 1)).from_buffer_copy(text + b"\0")
        assert lib.memo_lengths_text_parse(C.addressof(buf), len(text), final, st, out.ctypes.data, 1, want, res) == 0
        assert res[1] == 0
        return out[:res[0]].tolist(), res[2]

    for offset in range(0, 140, 1):
        for digits in range(1, 11):
            for sep in (b" ", b"\n", b"\t "):
                number = ("9" * digits if digits < 10 else "2147483647").encode()
                body = b" " * offset + number + sep + b"12 345 6 " * 20
                for tail in (b"", b"7", b"\n"):
                    text = body + tail
                    want = [int(x) for x in text.split()]
                    assert parse(text) == (want, len(text)), (offset, digits, sep, tail)
                    got, used = parse(text, final=0)
                    assert (got, used) == ((want[:-1], len(text) - 1) if tail == b"7" else (want, len(text)))
                    for budget in (1, 2, 5):
                        assert parse(text, want=budget)[0] == want[:budget]


def test_lengths_stream_blocks_threads_and_errors(tmp_path):
    """iter_lengths_columns: the same matrix whatever the block size, read buffer, tile and
    thread count; ragged, invalid and out-of-range files raise; a closed stream leaves no file open."""
    from memo_b200 import io
    from memo_b200._lib import MemoError
    rng = np.random.default_rng(77)
    L, C = 5000, 11
    vals = rng.integers(0, 3000, size=(L, C)).astype(np.int32)
    vals[rng.random((L, C)) < 0.01] = 2 ** 31 - 1
    paths = _write_lengths(tmp_path, vals)
    with open(paths[3], "rb+") as fh:                              # no newline at the end of one file
        fh.seek(-1, os.SEEK_END)
        fh.truncate()
    for kw in (dict(), dict(block_rows=7), dict(block_rows=1000, read_bytes=64, tile_rows=13, threads=3),
               dict(block_rows=L, threads=1), dict(block_rows=L - 1, read_bytes=4096, threads=32)):
        got = np.concatenate([b.copy() for b in io.iter_lengths_columns(paths, **kw)])
        assert got.dtype == np.int32 and np.array_equal(got, vals), kw
    assert np.array_equal(io.read_lengths_columns(paths), vals)
    empty = tmp_path / "empty.lengths"
    empty.write_text(">only a header\n\n  \n")
    assert io.read_lengths_columns([str(empty)]).shape == (0, 1)
    assert list(io.iter_lengths_columns([str(empty), str(empty)])) == []
    short = tmp_path / "short.lengths"
    short.write_text(">r\n" + " ".join(map(str, vals[:L - 1, 0])) + "\n")
    for pair in ([paths[0], str(short)], [str(short), paths[0]]):
        with pytest.raises(MemoError, match="expected the same number"):
            list(io.iter_lengths_columns(pair, block_rows=512))
    for text, exc in (("1 2 x 4\n", ValueError), ("1 -2 3\n", ValueError), ("5 >6\n", ValueError),
                      ("1 2147483648 3\n", MemoError), ("1 " + "9" * 80 + "\n", MemoError)):
        bad = tmp_path / "bad.lengths"
        bad.write_text(text)
        with pytest.raises(exc):
            io.read_lengths_columns([str(bad)])
    with pytest.raises(ValueError):                                # a "number" longer than the read buffer
        bad.write_text("7 " + "1" * 300 + " 8\n")
        list(io.iter_lengths_columns([str(bad)], read_bytes=128))
    def open_lengths_files():
        out = []
        for fd in os.listdir("/proc/self/fd"):
            try:
                out.append(os.readlink(f"/proc/self/fd/{fd}"))
            except OSError:
                pass
        return [t for t in out if t.startswith(str(tmp_path))]

    it = io.iter_lengths_columns(paths, block_rows=100)
    next(it)
    assert len(open_lengths_files()) == C
    it.close()
    assert open_lengths_files() == []


def test_read_int_text_reads_a_conservation_vector(tmp_path):
    """The input of `memo view` (one conservation value per line, memo_query.py:70) as
    plot_conservation.py:40-49 reads it; what int() rejects raises ValueError."""
    from memo_b200 import io
    rng = np.random.default_rng(5)
    vec = rng.integers(0, 70000, 100_001)
    p = tmp_path / "cons.txt"
    for text in ("\n".join(map(str, vec)) + "\n", "\n".join(map(str, vec)), " \n".join(map(str, vec)) + "\r\n"):
        p.write_text(text)
        got = io.read_int_text(str(p))
        assert got.dtype == np.int32 and np.array_equal(got, vec)
    p.write_text("")
    assert io.read_int_text(str(p)).size == 0
    p.write_text("7\n")
    assert io.read_int_text(str(p)).tolist() == [7]
    for bad in ("1\n-2\n", "1\nx\n", "1\n>2\n", ">h\n1\n", "99999999999\n", "1.0\n"):
        p.write_text(bad)
        with pytest.raises(ValueError):
            io.read_int_text(str(p))


def test_parquet_compress_bed_cli(example_golden, tmp_path):
    bed = tmp_path / "x.bed"
    bed.write_text(example_golden["cons_bed"])
    out = tmp_path / "x.parquet"
    r = subprocess.run([sys.executable, "-m", "memo_b200.parquet_compress_bed", "-f", str(bed), "-o", str(out)],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "DONE index compression" in r.stdout
    t = pq.read_table(str(out))
    assert [f.name for f in t.schema] == ["f0", "f1", "f2", "f3"]
    assert [str(f.type) for f in t.schema] == ["string", "int64", "int64", "int64"]
    assert t.schema.metadata is None
    assert pq.ParquetFile(str(out)).metadata.row_group(0).column(1).compression == "ZSTD"
    want = [l.split("\t") for l in example_golden["cons_bed"].splitlines()]
    assert t.column("f0").to_pylist() == [w[0] for w in want]
    for j in (1, 2, 3):
        assert t.column(f"f{j}").to_pylist() == [int(w[j]) for w in want]
    # the reader keeps the live predicate of memo_query.py:25-27
    from memo_b200 import io
    f1, f2, f3 = io.read_index_rows(str(out), "ref_1", 4, 9)
    assert f1.tolist() == [int(w[1]) for w in want if 4 < int(w[1]) < 9]


def test_dap_to_bed_argument_checks(tmp_path):
    from memo_b200 import dap_to_bed as d
    fai = tmp_path / "p.fai"; fai.write_text("a\t3\n")
    dap = tmp_path / "dap.txt"; dap.write_text("0 1\n")
    ok = d.parse_arguments(["--mem", "--overlap", "--order", "--fai", str(fai), "--dap", str(dap)])
    d.check_args(ok)
    assert ok.sort_lcps and ok.print_overlaps
    with pytest.raises(Exception, match="fai file does not exist"):
        d.check_args(d.parse_arguments(["--mem", "--fai", "nope.fai", "--dap", str(dap)]))
    with pytest.raises(Exception, match="incorrect file extension"):
        d.check_args(d.parse_arguments(["--mem", "--fai", str(dap), "--dap", str(dap)]))
    with pytest.raises(Exception, match="Either print MSs or MEMs"):
        d.check_args(d.parse_arguments(["--fai", str(fai), "--dap", str(dap)]))
    with pytest.raises(Exception, match="Either print MSs or MEMs"):
        d.check_args(d.parse_arguments(["--ms", "--mem", "--fai", str(fai), "--dap", str(dap)]))
    with pytest.raises(Exception, match="Can only print overlaps"):
        d.check_args(d.parse_arguments(["--ms", "--overlap", "--fai", str(fai), "--dap", str(dap)]))


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from memo_b200 import api, _lib
    with pytest.raises(_lib.MemoError):
        api.IndexBuilder()
    with pytest.raises(_lib.MemoError):
        api.synth_dap(100, 3, 1)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "memo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "oracle/" not in src.replace("oracle/memo_oracle.py:synth_dap", ""), f


def test_shard_plans_cover_everything():
    from memo_b200 import shard
    rng = np.random.default_rng(11)
    for _ in range(100):
        lens = [int(x) for x in rng.integers(1, 50, int(rng.integers(1, 5)))]
        recs = [(f"r{i}", l) for i, l in enumerate(lens)]
        total = sum(lens)
        world = int(rng.integers(1, 9))
        covered = 0
        for rank in range(world):
            p = shard.plan_shard(recs, total, world, rank, halo_right=7)
            assert p.pos_lo == covered
            covered = p.pos_hi
            owned = p.segs[:p.n_owned]
            assert sum(s.n_rows for s in owned) == p.pos_hi - p.pos_lo
            for s in p.segs:
                assert s.row_begin >= (0 if s.flags & 1 else 1)
                assert s.row_begin + s.n_rows <= p.buf_hi - p.buf_lo
        assert covered == total


def test_sharded_build_equals_whole_on_oracle():
    """Shard plans fed to the C oracle (same run semantics as the device ABI):
    concatenating the owned rows of every rank reproduces the unsharded index."""
    from memo_b200 import shard
    from oracle import c_oracle as co
    lens = [3000, 1, 4500, 2500]
    recs = [(f"c{i}", l) for i, l in enumerate(lens)]
    C = 6
    vals = np.concatenate([mo.synth_dap(n, C, seed=50 + i, dense=True) for i, n in enumerate(lens)])
    whole = co.index_build(vals, recs, True)
    for world in (1, 2, 3, 4, 8):
        parts = []
        for rank in range(world):
            p = shard.plan_shard(recs, len(vals), world, rank, halo_right=40)
            segs = [co.Seg(s.row_begin, s.n_rows, s.pos0, s.rec_len, s.rec_id, s.flags)
                    for s in p.segs[:p.n_owned]]
            parts.append(co.index_build(vals[p.buf_lo:p.buf_hi], recs, True, segs=segs))
        for j in range(4):
            assert np.array_equal(np.concatenate([q[j] for q in parts]), whole[j]), world


def test_query_rows_for_range_on_oracle():
    """a position range answered from its own rows plus a right halo of
    k_max - 1 positions equals the slice of the whole-window answer (oracle only)."""
    from memo_b200 import shard
    L, C = 3000, 7
    vals = mo.synth_dap(L, C, seed=5, dense=True)
    _, s, e, c = mo.index_build(vals, [("chrH", L)], True)
    ks = (3, 31, 64)
    for world in (2, 3, 5):
        for k in ks:
            whole = mo.query(s, e, c, 0, L, k, C + 1, False)
            parts = []
            for rank in range(world):
                lo, hi = shard.shard_range(L, world, rank)
                a, b = shard.query_rows_for_range(s, lo, hi, max(ks))
                parts.append(mo.query(s[a:b], e[a:b], c[a:b], lo, hi, k, C + 1, False))
            assert np.array_equal(np.concatenate(parts), whole), (world, k)


def test_sharded_build_partial_dap_keeps_chr_end_rows():
    """A DAP that stops inside the last record still gets that record's chr-end rows
    (src/dap_to_bed.py:133-134 runs after the last row whatever the .fai length)."""
    from memo_b200 import shard
    from oracle import c_oracle as co
    recs = [("a", 100), ("b", 50)]
    C = 4
    vals = np.concatenate([mo.synth_dap(100, C, seed=1, dense=True), mo.synth_dap(50, C, seed=2, dense=True)[:20]])
    whole = co.index_build(vals, recs, True)
    assert (whole[1] == 50).any()
    for world in (1, 2, 3):
        parts = []
        for rank in range(world):
            p = shard.plan_shard(recs, len(vals), world, rank)
            segs = [co.Seg(s.row_begin, s.n_rows, s.pos0, s.rec_len, s.rec_id, s.flags)
                    for s in p.segs[:p.n_owned]]
            parts.append(co.index_build(vals[p.buf_lo:p.buf_hi], recs, True, segs=segs))
        for j in range(4):
            assert np.array_equal(np.concatenate([q[j] for q in parts]), whole[j]), world


def test_dap_text_empty_field_raises_like_int(tmp_path):
    """An empty field (double space / ragged `paste` output) is int('') in the reference
    (src/dap_to_bed.py:87): ValueError, not a silent INT32_MIN."""
    from memo_b200 import io
    p = tmp_path / "dap.txt"
    p.write_text("0 3 5\n1  5\n2 4 4\n")
    with pytest.raises(ValueError):
        io.read_dap_text(str(p))
    p.write_text("0 3 5\n1 2 \n2 4 4\n")
    with pytest.raises(ValueError):
        io.read_dap_text(str(p))


def test_parquet_row_groups_prune_the_query_read(tmp_path):
    """the Parquet index is written in row groups cut at record changes with
    min/max statistics (IndexParquetWriter) and read_index_rows reads only the groups a window
    touches; rows, schema and codec are what parquet_compress_bed.py writes."""
    from memo_b200 import io
    rng = np.random.default_rng(0)
    recs = ["chrA", "chrB", "chrC"]
    cols = []
    for r, n in zip(recs, (50000, 300, 70000)):
        f1 = np.sort(rng.integers(1, 10 ** 6, n))
        cols.append((np.full(n, recs.index(r)), f1, f1 + rng.integers(0, 500, n), rng.integers(1, 10, n)))
    rec, f1, f2, f3 = (np.concatenate([c[i] for c in cols]) for i in range(4))
    table = io.index_table([(r, 0) for r in recs], rec, f1, f2, f3)
    path = tmp_path / "i.parquet"
    with io.IndexParquetWriter(str(path), rows_per_group=8000) as w:
        for a in range(0, table.num_rows, 3333):
            w.write(table.slice(a, 3333))
    back = pq.read_table(path)
    assert back.equals(table) and back.schema.equals(io.INDEX_SCHEMA) and back.schema.metadata is None
    md = pq.ParquetFile(path).metadata
    assert md.num_row_groups >= 15 and md.row_group(0).column(0).compression == "ZSTD"
    names = np.array(recs)[rec]
    for name, a, b, most in (("chrA", 1000, 60000, 2), ("chrB", 0, 10 ** 6, 1), ("chrC", 500000, 500100, 2),
                             ("nochr", 0, 5, 0)):
        st = {}
        g = io.read_index_rows(str(path), name, a, b, stats=st)
        sel = (names == name) & (f1 > a) & (f1 < b)
        assert np.array_equal(g[0], f1[sel]) and np.array_equal(g[1], f2[sel]) and np.array_equal(g[2], f3[sel])
        assert st["row_groups_read"] <= most < st["row_groups"], (name, st)
    # a file written by the reference's writer settings (one big group) still reads correctly
    pq.write_table(table, tmp_path / "plain.parquet", compression="ZSTD")
    g = io.read_index_rows(str(tmp_path / "plain.parquet"), "chrC", 500000, 500100)
    sel = (names == "chrC") & (f1 > 500000) & (f1 < 500100)
    assert np.array_equal(g[0], f1[sel])


def test_parquet_compress_bed_empty_bed_raises_like_the_reference(tmp_path):
    """An empty BED makes the reference crash (pyarrow: empty CSV; writer None at :39)."""
    from memo_b200 import parquet_compress_bed
    bed = tmp_path / "e.bed"
    bed.write_text("")
    with pytest.raises(Exception):
        parquet_compress_bed.bed_to_parquet(str(bed), str(tmp_path / "e.parquet"))


def test_parquet_compress_bed_blocks_are_invisible(tmp_path):
    """bed_to_parquet parses the BED in blocks on a helper thread: the table is the same for any
    block size (lines cut by a block, a last line without line end, one block, blocks smaller than
    a line), malformed rows raise in the caller's thread, and nothing is left running."""
    import threading
    from memo_b200 import io, parquet_compress_bed as pcb
    rng = np.random.default_rng(11)
    lines = []
    for name, n in (("chr1", 4000), ("a_rather_long_record_name_" * 4, 50), ("chrX", 2500)):
        f1 = np.sort(rng.integers(0, 10 ** 9, n))
        lines += [f"{name}\t{a}\t{a + d}\t{o}" for a, d, o in zip(f1, rng.integers(0, 10 ** 5, n), rng.integers(1, 94, n))]
    text = "\n".join(lines)                                        # no line end after the last row
    bed = tmp_path / "x.bed"
    bed.write_text(text)
    want = [ln.split("\t") for ln in lines]
    for block in (500_000_000, 4096, 100, 7):
        out = tmp_path / f"x{block}.parquet"
        pcb.bed_to_parquet(str(bed), str(out), block_size=block, rows_per_group=1000)
        t = pq.read_table(out)
        assert t.schema.equals(io.INDEX_SCHEMA) and t.schema.metadata is None
        assert t.column("f0").to_pylist() == [w[0] for w in want], block
        for j in (1, 2, 3):
            assert t.column(f"f{j}").to_pylist() == [int(w[j]) for w in want], block
    bad = tmp_path / "bad.bed"
    bad.write_text("\n".join(lines[:2000]) + "\nchr1\tx\t5\t1\n" + "\n".join(lines[2000:]) + "\n")
    with pytest.raises(Exception, match="(?i)conversion|invalid|CSV"):
        pcb.bed_to_parquet(str(bad), str(tmp_path / "bad.parquet"), block_size=4096)
    assert not [th for th in threading.enumerate() if th.name == "memo-bed-parse"]


def test_stream_kernels_keep_their_register_budget():
    """Occupancy is part of the design: four CTAs of six warps of the strip
    kernel per SM need <= 85 registers per thread (ptxas settles on 80; `__launch_bounds__(256, 1)`
    instead of `__launch_bounds__(256)` once let it take 105 and halved the occupancy)."""
    import shutil
    from memo_b200 import _build
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    _build.build()
    out = subprocess.run([exe, "-res-usage", os.path.join(_build.CSRC, "index_wide.o")], capture_output=True,
                         text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    wide = {k: v for k, v in regs.items() if "wide_kernelILi3E" in k}
    assert len(wide) == 2 and all(v <= 85 for v in wide.values()), wide


def test_pread_into_reads_ranges_with_and_without_threads(tmp_path):
    """io._pread_into (the dap.txt block reader): any offset / length, the end of the file inside
    a block, one thread and several."""
    import os
    from memo_b200 import io
    data = np.random.default_rng(0).integers(0, 256, 3_000_001, dtype=np.uint8).tobytes()
    path = tmp_path / "blob"
    path.write_bytes(data)
    fd = os.open(path, os.O_RDONLY)
    try:
        for threads in (1, 4):
            for off, n in ((0, len(data)), (7, 1 << 20), (123, 2_500_000), (2_900_000, 500_000),
                           (len(data), 100), (0, 0)):
                buf = bytearray(n)
                got = io._pread_into(fd, memoryview(buf), off, piece=1 << 18, threads=threads)
                want = data[off:off + n]
                assert got == len(want) and bytes(buf[:got]) == want, (threads, off, n)
    finally:
        os.close(fd)
