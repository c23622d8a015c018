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


def bed_to_parquet(bed_path, parquet_path, block_size=500_000_000, codec="ZSTD", rows_per_group=1 << 21):
    """Stream the BED in `block_size`-byte blocks (src/parquet_compress_bed.py:16-38); the rows
    go out in row groups of about `rows_per_group` rows cut at record changes, with min/max
    statistics, so that a query reads only the groups its window touches (the reference writes
    one group per 500 MB block of text: every query scans all of it)."""
    reader = pacsv.open_csv(
        bed_path,
        read_options=pacsv.ReadOptions(column_names=INDEX_SCHEMA.names, block_size=int(block_size)),
        parse_options=pacsv.ParseOptions(delimiter="\t"),
        convert_options=pacsv.ConvertOptions(column_types=INDEX_SCHEMA),
    )
    with reader, IndexParquetWriter(parquet_path, codec=codec, rows_per_group=rows_per_group) as sink:
        for batch in reader:
            sink.write(pa.Table.from_batches([batch]))


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
