"""Host-side file formats of the hot path: dap.txt ingest, BED / Parquet index
writers and readers.  Format conversion only -- no index arithmetic here."""
from __future__ import annotations

import re
import sys
import warnings
from typing import Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.csv as pacsv
import pyarrow.dataset as pads
import pyarrow.parquet as pq

from ._lib import MemoError

_NOT_LENGTHS = re.compile(rb"[^0-9\s]")

INDEX_SCHEMA = pa.schema([("f0", pa.utf8()), ("f1", pa.int64()), ("f2", pa.int64()), ("f3", pa.int64())])


def read_dap_text(path) -> Tuple[int, np.ndarray]:
    """dap.txt (index.sh:83: `pos v1 ... vC`, single spaces) -> (first position,
    int32 [L, C] matrix).  The device path needs consecutive positions (what
    `nl -v0` produces); anything else is rejected loudly."""
    table = pacsv.read_csv(
        path,
        read_options=pacsv.ReadOptions(autogenerate_column_names=True),
        parse_options=pacsv.ParseOptions(delimiter=" "),
    )
    if table.num_columns < 2:
        raise MemoError("dap.txt needs a position column and at least one genome column")
    for i, col in enumerate(table.columns):
        if not pa.types.is_integer(col.type) or col.null_count:
            # int() in the reference (an empty field -- double space, ragged `paste` output --
            # parses as null here and raises there: src/dap_to_bed.py:87)
            raise ValueError(f"invalid literal for int() in DAP column {i}")
    L = table.num_rows
    if L == 0:
        return 0, np.zeros((0, table.num_columns - 1), dtype=np.int32)
    pos = table.column(0).to_numpy()
    if L > 1 and not (np.diff(pos) == 1).all():
        raise MemoError("dap.txt positions are not consecutive (expected `nl -v0` numbering)")
    C = table.num_columns - 1
    out = np.empty((L, C), dtype=np.int32)
    for j in range(C):
        col = table.column(j + 1)
        mm = pc.min_max(col).as_py()
        if mm["min"] < 0 or mm["max"] > 2**31 - 1:
            raise MemoError("DAP lengths must be in [0, 2^31)")
        out[:, j] = col.to_numpy()
    return int(pos[0]), out


def read_lengths_columns(paths: Sequence[str], threads: int = 8) -> np.ndarray:
    """Per-genome MONI `*.lengths` (or the `*.lengths.vert` index.sh:79 makes of them) ->
    int32 [L, C] DAP matrix, one column per file in the order given (genome_list.txt
    order minus the pivot).  Replaces index.sh:79-83 (`grep -v '^>' | tr ' ' '\n'`,
    `paste | nl`) and the text re-parse of dap.txt (src/dap_to_bed.py:87): header lines
    start with '>', every other line holds whitespace-separated lengths; row i of the
    result is pivot position i of the concatenated records."""
    from concurrent.futures import ThreadPoolExecutor

    def one(path):
        with open(path, "rb") as fh:
            data = fh.read()
        if b">" in data:
            data = b"\n".join(ln for ln in data.split(b"\n") if not ln.startswith(b">"))
        if _NOT_LENGTHS.search(data):                            # int() in the reference
            raise ValueError(f"invalid literal for int() in {path}")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", DeprecationWarning)
            col = np.fromstring(data.decode("ascii"), dtype=np.int64, sep=" ")     # any whitespace separates
        if col.size and (col.min() < 0 or col.max() > 2**31 - 1):
            raise MemoError("DAP lengths must be in [0, 2^31)")
        return col.astype(np.int32)

    if not paths:
        raise MemoError("at least one .lengths file is needed")
    with ThreadPoolExecutor(max(1, min(threads, len(paths)))) as ex:
        cols = list(ex.map(one, paths))
    L = cols[0].size
    for path, col in zip(paths, cols):
        if col.size != L:
            raise MemoError(f"{path}: {col.size} lengths, expected {L} (one per pivot position)")
    out = np.empty((L, len(cols)), dtype=np.int32)
    for j, col in enumerate(cols):
        out[:, j] = col
    return out


def index_table(records: Sequence[Tuple[str, int]], rec_idx, start, end, order) -> pa.Table:
    """Arrow table with the index schema (f0 string, f1..f3 int64), f0
    dictionary-free so that it equals what parquet_compress_bed.py reads back."""
    names = pa.array([r[0] for r in records], type=pa.utf8())
    f0 = pc.take(names, pa.array(np.asarray(rec_idx, dtype=np.int64)))
    return pa.table([f0, pa.array(np.asarray(start, dtype=np.int64)),
                     pa.array(np.asarray(end, dtype=np.int64)),
                     pa.array(np.asarray(order, dtype=np.int64))], schema=INDEX_SCHEMA)


def write_bed(table: pa.Table, sink=None) -> None:
    """BED payload as dap_to_bed.py prints it: tab separated, no header."""
    sink = sys.stdout.buffer if sink is None else sink
    if table.num_rows == 0:
        return
    pacsv.write_csv(table, sink, write_options=pacsv.WriteOptions(
        include_header=False, delimiter="\t", quoting_style="none"))


def write_parquet(table: pa.Table, path, codec: str = "ZSTD") -> None:
    pq.write_table(table, path, compression=codec)


def read_index_rows(path, record: str, f1_gt: int, f1_lt: int):
    """Index rows of `record` with f1_gt < f1 < f1_lt (the live predicate of
    memo_query.py:25-27; the other predicate's rows can never paint)."""
    dataset = pads.dataset(path, format="parquet")
    flt = (pads.field("f0") == record) & (pads.field("f1") > f1_gt) & (pads.field("f1") < f1_lt)
    t = dataset.to_table(filter=flt, columns=["f1", "f2", "f3"])
    return (t.column("f1").to_numpy(), t.column("f2").to_numpy(), t.column("f3").to_numpy())
