#!/usr/bin/env python3
"""Drop-in for the reference's src/parquet_compress_bed.py (BED -> Parquet).

Pure format conversion on the host with pyarrow (north_star: "parquet/BED index
format unchanged"); no device work.  Same argv (src/parquet_compress_bed.py:
51-60), same output schema f0:string, f1..f3:int64, ZSTD, row order = BED order.
"""
import argparse

import pyarrow as pa
import pyarrow.csv as pacsv

from .io import INDEX_SCHEMA, IndexParquetWriter


SEGMENT_BYTES = 128 << 20          # BED text parsed per step (an upper bound on -b)


def _bed_segments(bed_path, seg_bytes):
    """The BED file as a sequence of byte blocks of about `seg_bytes`, each ending at a line end."""
    with open(bed_path, "rb") as fh:
        tail = b""
        while True:
            data = fh.read(seg_bytes)
            if not data:
                if tail:
                    yield tail
                return
            cut = data.rfind(b"\n") + 1
            if cut == 0:                      # no line end in this piece: the line goes on
                tail += data
                continue
            yield tail + data[:cut] if tail else (data if cut == len(data) else data[:cut])
            tail = data[cut:]


def bed_to_parquet(bed_path, parquet_path, block_size=500_000_000, codec="ZSTD", rows_per_group=1 << 21):
    """BED text -> Parquet index, block by block (src/parquet_compress_bed.py:16-38: blocks of
    `block_size` bytes of text, one table each, appended to one ParquetWriter).  Here a block is at
    most 128 MB, parsed by pyarrow's multi-threaded CSV reader on a helper thread while the previous
    block's rows are encoded and written (the reference's streaming reader parses on one thread and
    alternates with the writer); the rows go out in row groups of about `rows_per_group` rows cut at
    record changes, with min/max statistics, so that a query reads only the groups its window
    touches (the reference writes one group per block of text: every query scans all of it)."""
    import os
    import queue
    import threading
    read_opts = dict(column_names=INDEX_SCHEMA.names, block_size=8 << 20)
    parse_opts = pacsv.ParseOptions(delimiter="\t")
    conv_opts = pacsv.ConvertOptions(column_types=INDEX_SCHEMA)
    if os.path.getsize(bed_path) == 0:        # pyarrow's own error, as in the reference ("Empty CSV file")
        pacsv.read_csv(bed_path, read_options=pacsv.ReadOptions(**read_opts), parse_options=parse_opts,
                       convert_options=conv_opts)
    tables = queue.Queue(maxsize=2)
    stop = threading.Event()

    def parse():
        try:
            for seg in _bed_segments(bed_path, max(1, min(int(block_size), SEGMENT_BYTES))):
                if stop.is_set():
                    return
                tables.put(pacsv.read_csv(pa.BufferReader(seg), read_options=pacsv.ReadOptions(**read_opts),
                                          parse_options=parse_opts, convert_options=conv_opts))
            tables.put(None)
        except BaseException as exc:          # handed to the writer's thread
            tables.put(exc)

    worker = threading.Thread(target=parse, name="memo-bed-parse", daemon=True)
    worker.start()
    try:
        with IndexParquetWriter(parquet_path, codec=codec, rows_per_group=rows_per_group) as sink:
            while True:
                item = tables.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                sink.write(item)
    finally:
        stop.set()
        while worker.is_alive():              # let a parser blocked on the full queue finish
            try:
                tables.get_nowait()
            except queue.Empty:
                pass
            worker.join(0.05)


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser(description="Converts input bed file to Parquet file.")
    ap.add_argument("-f", "--file", dest="file", required=True, help="bed file")
    ap.add_argument("-o", "--output", dest="output", default=None, help="output parquet [FILE.parquet]")
    ap.add_argument("-b", "--block_size", dest="block_size", default=500_000_000,
                    help="block size in bytes [500_000_000]")
    ap.add_argument("-c", "--codec", dest="codec", default="ZSTD", help="compression codec [ZSTD]")
    ap.add_argument("-a", "--all", dest="compress_all_at_once", action="store_true", default=False,
                    help="convert in one block")
    return ap.parse_args(argv)


def main(args):
    out_path = args.output if args.output else args.file.rstrip(".bed") + ".parquet"   # same quirk
    print("Input bed:", args.file)
    print("Output parquet:", out_path)
    print("Code:", args.codec)
    if args.compress_all_at_once:
        print("Compressing bed file all at once.")
        bed_to_parquet(args.file, out_path, block_size=2**31 - 1, codec=args.codec)
    else:
        print("Block size (bytes):", args.block_size)
        bed_to_parquet(args.file, out_path, block_size=int(args.block_size), codec=args.codec)
    print("DONE index compression")


if __name__ == "__main__":
    main(parse_arguments())
