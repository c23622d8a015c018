"""Host-side file formats of the hot path: dap.txt ingest, BED / Parquet index
writers and readers.  Format conversion only -- no index arithmetic here."""
from __future__ import annotations

import io as _pyio
import os as _os
import re
import sys
import warnings
from typing import Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.csv as pacsv
import pyarrow.parquet as pq

from ._lib import MemoError

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


class _RangeFile(_pyio.RawIOBase):
    """Bytes [lo, hi) of a file as a read-only stream (a rank's share of dap.txt)."""

    def __init__(self, path, lo, hi):
        super().__init__()
        self.fh, self.left = open(path, "rb"), hi - lo
        self.fh.seek(lo)

    def readable(self):
        return True

    def readinto(self, b):
        n = min(len(b), self.left)
        if n <= 0:
            return 0
        got = self.fh.readinto(memoryview(b)[:n])
        self.left -= got
        return got

    def close(self):
        self.fh.close()
        super().close()


def split_text_rows(path, world: int, rank: int):
    """A rank's share of a text file of rows, cut at line starts near equal byte offsets:
    (byte_lo, byte_hi, the line before byte_lo or None).  Shares of all ranks tile the file."""
    size = _os.path.getsize(path)

    def line_start(target):
        if target <= 0:
            return 0
        if target >= size:
            return size
        with open(path, "rb") as fh:
            fh.seek(target - 1)
            while True:                                    # first line start at or after target
                chunk = fh.read(1 << 16)
                if not chunk:
                    return size
                i = chunk.find(b"\n")
                if i >= 0:
                    return fh.tell() - len(chunk) + i + 1

    lo, hi = line_start(size * rank // world), line_start(size * (rank + 1) // world)
    prev = None
    if lo > 0:
        with open(path, "rb") as fh:
            back = min(lo, 1 << 20)
            fh.seek(lo - back)
            tail = fh.read(back)
            prev = tail[:-1].rsplit(b"\n", 1)[-1]
    return lo, hi, prev


def iter_dap_text(path, block_bytes: int = 64 << 20, byte_range=None):
    """dap.txt streamed in blocks (the reference streams it row by row, src/dap_to_bed.py:14-18):
    yields (first position, int32 [n, C]) for consecutive runs of rows; memory use is
    O(block_bytes).  Same validation as read_dap_text.  byte_range = (lo, hi): only the lines in
    those bytes (a rank's share, split_text_rows)."""
    lo, hi = byte_range if byte_range is not None else (0, _os.path.getsize(path))
    if hi <= lo:
        return
    with open(path, "rb") as fh:
        fh.seek(lo)
        first = fh.readline()
    n_fields = len(first.split(b" ")) if first.strip() else 0
    if n_fields == 0:
        return
    if n_fields < 2:
        raise MemoError("dap.txt needs a position column and at least one genome column")
    names = [f"f{i}" for i in range(n_fields)]
    try:
        reader = pacsv.open_csv(
            pa.PythonFile(_RangeFile(path, lo, hi), mode="r"),
            read_options=pacsv.ReadOptions(column_names=names, block_size=block_bytes),
            parse_options=pacsv.ParseOptions(delimiter=" "),
            convert_options=pacsv.ConvertOptions(column_types={n: pa.int64() for n in names}),
        )
        expect = None
        for batch in reader:
            n = batch.num_rows
            if n == 0:
                continue
            cols = batch.columns
            for i, col in enumerate(cols):
                if col.null_count:
                    raise ValueError(f"invalid literal for int() in DAP column {i}")
            pos = cols[0].to_numpy(zero_copy_only=False)
            if (n > 1 and not (np.diff(pos) == 1).all()) or (expect is not None and int(pos[0]) != expect):
                raise MemoError("dap.txt positions are not consecutive (expected `nl -v0` numbering)")
            expect = int(pos[-1]) + 1
            out = np.empty((n, n_fields - 1), dtype=np.int32)
            for j in range(1, n_fields):
                a = cols[j].to_numpy(zero_copy_only=False)
                if a.min() < 0 or a.max() > 2**31 - 1:
                    raise MemoError("DAP lengths must be in [0, 2^31)")
                out[:, j - 1] = a
            yield int(pos[0]), out
    except pa.ArrowInvalid as e:                  # int() in the reference (src/dap_to_bed.py:87)
        raise ValueError(f"invalid literal for int() in dap.txt: {e}") from None


_READ_POOL = None


def _pread_into(fd: int, mv: memoryview, offset: int, piece: int = 8 << 20, threads: int = 0) -> int:
    """Fill `mv` from file offset `offset`, optionally with several threads (os.preadv releases
    the GIL; one thread copies out of the page cache at a few GB/s, which is most of the text
    path's time).  threads = 0: MEMO_TEXT_READ_THREADS (default 4: 4.9 -> 8.9 GB/s of text on the B200 box).  Returns the number of bytes
    read: less than len(mv) only at the end of the file."""
    global _READ_POOL
    n = len(mv)
    if threads <= 0:
        threads = int(_os.environ.get("MEMO_TEXT_READ_THREADS", "4"))
    if n <= piece or threads <= 1:
        got = 0
        while got < n:
            r = _os.preadv(fd, [mv[got:]], offset + got)
            if r <= 0:
                break
            got += r
        return got
    if _READ_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _READ_POOL = ThreadPoolExecutor(threads)

    def one(a):
        b, got = min(n, a + piece), 0
        while a + got < b:
            r = _os.preadv(fd, [mv[a + got:b]], offset + a + got)
            if r <= 0:
                break
            got += r
        return got
    counts = list(_READ_POOL.map(one, range(0, n, piece)))
    total = 0
    for c, a in zip(counts, range(0, n, piece)):
        total += c
        if c < min(piece, n - a):                    # end of file inside this piece
            break
    return total


def iter_dap_text_device(path, device=None, block_bytes: int = 128 << 20, byte_range=None):
    """dap.txt parsed ON THE DEVICE: the file is read in blocks of whole lines into pinned
    memory, the bytes go to the GPU and memo_dap_text_parse turns them into int32 rows (the
    host never looks at a digit).  Yields (first position, device int32 [n, C]); a block is a
    view of the parser's buffer, valid until the next one is requested.  byte_range = (lo, hi):
    a rank's share of the file (split_text_rows)."""
    import torch
    from . import api
    lo, hi = byte_range if byte_range is not None else (0, _os.path.getsize(path))
    if hi <= lo:
        return
    with open(path, "rb") as fh:
        fh.seek(lo)
        first = fh.readline()
        n_fields = len(first.split(b" ")) if first.strip() else 0
        if n_fields == 0:
            return
        if n_fields < 2:
            raise MemoError("dap.txt needs a position column and at least one genome column")
        block_bytes = max(int(block_bytes), 4 * len(first) + 64)
        parser = api.DapTextParser(n_fields - 1, block_bytes, device)
        # two pinned blocks: the next one is read (by a helper thread; the read itself fans out,
        # _pread_into) while the current one is parsed, built and written by the caller
        pins = [torch.empty(block_bytes + 64, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        views = [p.numpy() for p in pins]
        from concurrent.futures import ThreadPoolExecutor
        ahead = ThreadPoolExecutor(1)
        fd = fh.fileno()

        def read(i, carry, off, want):
            return _pread_into(fd, memoryview(views[i])[carry:carry + want], off) if want > 0 else 0

        left, carry, off, cur = hi - lo, 0, lo, 0      # bytes of the share not read yet; bytes carried over
        fut = ahead.submit(read, cur, carry, off, min(left, block_bytes - carry))
        try:
            while True:
                got = fut.result()
                off += got
                left -= got
                n = carry + got
                if got == 0 and left > 0:
                    left = 0                           # (file shorter than expected)
                view = views[cur]
                if n == 0:
                    break
                if left == 0:
                    if view[n - 1] != 10:              # last line of the share without its newline
                        view[n] = 10
                        n += 1
                    cut = n
                else:
                    # the block ends with its last whole line; the rest opens the next block
                    tail = view[max(0, n - (1 << 16)):n]
                    nl = np.flatnonzero(tail == 10)
                    if nl.size == 0:
                        nl = np.flatnonzero(view[:n] == 10)
                        if nl.size == 0:
                            raise MemoError("dap.txt: a line longer than the text block")
                        cut = int(nl[-1]) + 1
                    else:
                        cut = max(0, n - (1 << 16)) + int(nl[-1]) + 1
                if cut == 0:
                    break
                head = bytes(view[:min(cut, 24)])
                try:
                    pos_first = int(head.split(b" ", 1)[0])
                except ValueError:
                    raise ValueError("invalid literal for int() in dap.txt") from None
                carry = n - cut
                more = left > 0 or carry > 0
                if more:
                    if carry:
                        views[1 - cur][:carry] = view[cut:n]
                    fut = ahead.submit(read, 1 - cur, carry, off, min(left, block_bytes - carry))
                rows = parser.parse(pins[cur], cut, pos_first)    # (synchronises: the pinned block is free again)
                if rows.shape[0]:
                    yield pos_first, rows
                if not more:
                    break
                cur = 1 - cur
        finally:
            ahead.shutdown(wait=True)                  # (a read still in flight targets the pinned blocks)


def iter_lengths_columns(paths: Sequence[str], block_rows: int = 1 << 18, read_bytes: int = 1 << 20,
                         threads: int = 0, tile_rows: int = 2048):
    """Per-genome MONI `*.lengths` / `*.lengths.vert` files streamed side by side: yields int32
    [n, C] blocks of consecutive pivot positions, column j from paths[j] (genome_list.txt order
    minus the pivot).  Replaces index.sh:79-83 (`grep -v '^>' | tr ' ' '\n' | grep .`,
    `paste | nl`) and the text re-parse of dap.txt (src/dap_to_bed.py:87): lines that start with
    '>' are headers, everything else is white-space separated lengths; row i of the stream is
    pivot position i of the concatenated records.  Memory is O(C * block): all files advance
    together.  The files are tokenized by `threads` host threads (default: MEMO_LENGTHS_THREADS
    or up to 16) inside libmemo_b200.so (memo_lengths_block_parse: host code, the GIL is
    released), each owning a group of neighbouring columns and walking the block in tiles of
    `tile_rows` rows; the next block is parsed while the caller works on the current one, so a
    yielded block is valid until the block after it is requested."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from . import _lib
    if not paths:
        raise MemoError("at least one .lengths file is needed")
    lib = _lib.load()
    n_cols = len(paths)
    if threads <= 0:
        threads = int(_os.environ.get("MEMO_LENGTHS_THREADS", 0)) or min(16, _os.cpu_count() or 1)
    threads = max(1, min(threads, n_cols))
    cuts = [k * n_cols // threads for k in range(threads + 1)]
    files = (_lib.LengthsFile * n_cols)()
    bufs_in = []                                                 # keeps the read buffers alive
    pool = ThreadPoolExecutor(threads)
    ahead = ThreadPoolExecutor(1)
    fut = None
    try:
        for j, p in enumerate(paths):
            files[j].fd = -1
        for j, p in enumerate(paths):
            buf = (C.c_uint8 * read_bytes)()
            bufs_in.append(buf)
            files[j].buf = C.cast(buf, C.c_void_p)
            files[j].cap = read_bytes
            files[j].fd = _os.open(p, _os.O_RDONLY)

        def fill(k, base):
            a, b = cuts[k], cuts[k + 1]
            group = C.cast(C.byref(files, a * C.sizeof(_lib.LengthsFile)), C.POINTER(_lib.LengthsFile))
            return lib.memo_lengths_block_parse(group, b - a, base + 4 * a, n_cols, block_rows, tile_rows)

        def produce(out):
            first = files[0].count
            base = out.ctypes.data
            rcs = list(pool.map(lambda k: fill(k, base), range(threads)))
            for j, p in enumerate(paths):
                err = files[j].error
                if err & 8:
                    raise MemoError("DAP lengths must be in [0, 2^31)")
                if err & 32:
                    raise OSError(f"reading {p} failed")
                if err:                                          # int() raises in the reference
                    raise ValueError(f"invalid literal for int() in {p}")
            if any(rcs):
                raise MemoError(f"memo_lengths_block_parse failed (codes {rcs})")
            for j, p in enumerate(paths):
                if files[j].count != files[0].count:
                    which = "fewer" if files[j].count < files[0].count else "more"
                    raise MemoError(f"{p}: {which} lengths than {paths[0]}, expected the same number "
                                    "(one per pivot position)")
            return files[0].count - first

        bufs = [np.empty((block_rows, n_cols), dtype=np.int32) for _ in range(2)]
        k = 0
        fut = ahead.submit(produce, bufs[0])
        while True:
            n, fut = fut.result(), None
            out = bufs[k]
            if n < block_rows:                                    # every file has ended
                if n:
                    yield out[:n]
                return
            k = 1 - k
            fut = ahead.submit(produce, bufs[k])
            yield out
    finally:
        if fut is not None:
            try:
                fut.result()
            except Exception:
                pass
        ahead.shutdown(wait=True)
        pool.shutdown(wait=True)
        for j in range(n_cols):
            if files[j].fd >= 0:
                _os.close(files[j].fd)
                files[j].fd = -1


def read_int_text(path) -> np.ndarray:
    """White-space separated non-negative decimal integers of a text file -> int32 vector (the
    conservation vector `memo query` writes, one value per line: src/plot_conservation.py:40-49
    reads it with int(line.strip()) per line).  Tokenized by the library's host tokenizer; anything
    int() would reject, and a '>' anywhere, raises ValueError."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    with open(path, "rb") as fh:
        data = fh.read()
    if b">" in data:
        raise ValueError(f"invalid literal for int() in {path}")
    out = np.empty((len(data) + 1) // 2, dtype=np.int32)          # a value and its separator take two bytes
    state, res = C.c_int32(0), (C.c_int64 * 3)()
    buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
    rc = lib.memo_lengths_text_parse(C.addressof(buf) if data else None, len(data), 1, state,
                                     out.ctypes.data, 1, out.size, res)
    if rc != 0:
        raise MemoError(f"memo_lengths_text_parse failed (code {rc})")
    if res[1]:
        raise ValueError(f"invalid literal for int() in {path}" if res[1] & 1 else f"value out of range in {path}")
    return out[:res[0]].copy() if res[0] * 4 < out.size else out[:res[0]]


def read_lengths_columns(paths: Sequence[str], threads: int = 0) -> np.ndarray:
    """The whole int32 [L, C] DAP matrix of per-genome MONI files (iter_lengths_columns in one piece)."""
    blocks = [b.copy() for b in iter_lengths_columns(paths, threads=threads)]
    if not blocks:
        return np.empty((0, len(paths)), dtype=np.int32)
    return blocks[0] if len(blocks) == 1 else np.concatenate(blocks)


def index_table(records: Sequence[Tuple[str, int]], rec_idx, start, end, order) -> pa.Table:
    """Arrow table with the index schema (f0 string, f1..f3 int64), f0
    dictionary-free so that it equals what parquet_compress_bed.py reads back."""
    names = pa.array([r[0] for r in records], type=pa.utf8())
    f0 = pc.take(names, pa.array(np.asarray(rec_idx, dtype=np.int64)))
    return pa.table([f0, pa.array(np.asarray(start, dtype=np.int64)),
                     pa.array(np.asarray(end, dtype=np.int64)),
                     pa.array(np.asarray(order, dtype=np.int64))], schema=INDEX_SCHEMA)


def index_batch(names: pa.Array, rec_counts, start, end, order) -> pa.Table:
    """Arrow table of one batch of index rows: rec_counts = [(record index, rows)] in row order,
    start / end / order int32 / uint32 / int32 numpy columns."""
    ids = np.repeat(np.array([r for r, _ in rec_counts], dtype=np.int64),
                    np.array([c for _, c in rec_counts], dtype=np.int64))
    return pa.table([pc.take(names, pa.array(ids)), pa.array(start.astype(np.int64)),
                     pa.array(end.astype(np.int64)), pa.array(order.astype(np.int64))], schema=INDEX_SCHEMA)


def write_bed(table: pa.Table, sink=None) -> None:
    """BED payload as dap_to_bed.py prints it: tab separated, no header."""
    sink = sys.stdout.buffer if sink is None else sink
    if table.num_rows == 0:
        return
    pacsv.write_csv(table, sink, write_options=pacsv.WriteOptions(
        include_header=False, delimiter="\t", quoting_style="none"))


class IndexParquetWriter:
    """Parquet index (schema f0:string, f1..f3:int64, ZSTD -- src/parquet_compress_bed.py:19-38)
    written incrementally with row groups sized for the query's predicates: a row group holds
    about `rows_per_group` rows and, where a record is large enough, rows of one record only,
    so that the f0 / f1 min-max statistics of a group bound a position range of one record and
    read_index_rows can skip every group a window does not touch (src/memo_query.py:25-27 scans
    the whole file).  Row order, schema and codec are the reference's; row-group boundaries are
    not observable by memo_query.py."""

    def __init__(self, path, codec: str = "ZSTD", rows_per_group: int = 1 << 21):
        self.rows_per_group = int(rows_per_group)
        self.writer = pq.ParquetWriter(path, INDEX_SCHEMA, compression=codec, write_statistics=True)
        self.pending = []                 # tables of the group being collected
        self.n_pending = 0
        self.last_f0 = None
        self.n_rows = 0

    def write(self, table: pa.Table) -> None:
        """Rows in index order (any batch size)."""
        if table.num_rows == 0:
            return
        if not table.schema.equals(INDEX_SCHEMA):
            table = table.cast(INDEX_SCHEMA)
        f0 = table.column("f0").combine_chunks()
        # cut the batch where the record changes
        change = pc.not_equal(f0.slice(1), f0.slice(0, len(f0) - 1)).to_numpy(zero_copy_only=False)
        cuts = [0] + (np.flatnonzero(change) + 1).tolist() + [table.num_rows]
        for a, b in zip(cuts[:-1], cuts[1:]):
            name = f0[a].as_py()
            # a new record starts a new group once the current one is worth a group of its own
            if self.last_f0 is not None and name != self.last_f0 and self.n_pending >= self.rows_per_group // 8:
                self._flush()
            self.last_f0 = name
            pos = a
            while pos < b:
                take = min(b - pos, self.rows_per_group - self.n_pending)
                self.pending.append(table.slice(pos, take))
                self.n_pending += take
                pos += take
                if self.n_pending >= self.rows_per_group:
                    self._flush()

    def _flush(self) -> None:
        if self.n_pending:
            t = pa.concat_tables(self.pending)
            self.writer.write_table(t, row_group_size=t.num_rows)
            self.n_rows += t.num_rows
        self.pending, self.n_pending = [], 0

    def close(self) -> None:
        self._flush()
        self.writer.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write_parquet(table: pa.Table, path, codec: str = "ZSTD") -> None:
    pq.write_table(table, path, compression=codec)


def read_index_rows(path, record: str, f1_gt: int, f1_lt: int, stats: dict = None):
    """Index rows of `record` with f1_gt < f1 < f1_lt (the live predicate of
    memo_query.py:25-27; the other predicate's rows can never paint).  Only the row
    groups whose f0 / f1 statistics admit such a row are read (IndexParquetWriter sizes the
    groups for that; a file without statistics is read whole); `stats` receives the counts."""
    pf = pq.ParquetFile(path)
    md = pf.metadata
    names = md.schema.names
    c0, c1 = names.index("f0"), names.index("f1")
    keep = []
    for g in range(md.num_row_groups):
        rg = md.row_group(g)
        s0, s1 = rg.column(c0).statistics, rg.column(c1).statistics
        ok = True
        if s0 is not None and s0.has_min_max:
            lo, hi = s0.min, s0.max
            lo = lo.decode() if isinstance(lo, bytes) else lo
            hi = hi.decode() if isinstance(hi, bytes) else hi
            ok = lo <= record <= hi
        if ok and s1 is not None and s1.has_min_max:
            ok = s1.max > f1_gt and s1.min < f1_lt
        if ok:
            keep.append(g)
    if stats is not None:
        stats.update(row_groups=md.num_row_groups, row_groups_read=len(keep))
    if not keep:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), z.copy()
    t = pf.read_row_groups(keep, columns=["f0", "f1", "f2", "f3"])
    f1 = t.column("f1")
    mask = pc.and_(pc.equal(t.column("f0"), record), pc.and_(pc.greater(f1, f1_gt), pc.less(f1, f1_lt)))
    t = t.filter(mask)
    return (t.column("f1").to_numpy(), t.column("f2").to_numpy(), t.column("f3").to_numpy())
