#!/usr/bin/env python3
"""Drop-in for the reference's src/dap_to_bed.py on the B200 device path.

Same argv (src/dap_to_bed.py:139-149), same validation (:151-171), BED payload
on stdout and nothing else (index.sh:93,102 redirect stdout into the .bed).

    python -m memo_b200.dap_to_bed --mem [--order] --overlap --fai P.fai --dap dap.txt > out.bed

Extension (not in the reference): `--lengths G2.lengths G3.lengths ...` instead of
`--dap` reads the per-genome MONI outputs directly (genome_list.txt order minus the
pivot) and skips index.sh:79-83's vertical files, `paste | nl` and the text re-parse.
"""
import argparse
import os
import sys


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser(description="Takes in .fai and full document array profile and "
                                 "converts to bed-style MEM intervals to stdout (B200 device path).")
    ap.add_argument("--fai", dest="fai_path", required=True, help="path to fai file")
    ap.add_argument("--dap", dest="dap_path", default=None, help="path to full document profile")
    ap.add_argument("--lengths", dest="lengths_paths", nargs="+", default=None,
                    help="extension: per-genome MONI .lengths files instead of --dap")
    ap.add_argument("--ms", dest="print_ms", action="store_true", default=False,
                    help="Extract matching statistics (either MSs or MEMs, not both).")
    ap.add_argument("--mem", dest="print_mems", action="store_true", default=False,
                    help="Extract MEMs (either MSs or MEMs, not both).")
    ap.add_argument("--overlap", dest="print_overlaps", action="store_true", default=False,
                    help="extract overlap MEMs (only with --mem).")
    ap.add_argument("--order", dest="sort_lcps", action="store_true", default=False,
                    help="sort LCP row to extract order MS/MEMs.")
    return ap.parse_args(argv)


def check_args(args):
    if not os.path.isfile(args.fai_path):
        raise Exception("The fai file does not exist.")
    if args.lengths_paths is None:
        if args.dap_path is None or not os.path.isfile(args.dap_path):
            raise Exception("The dap file does not exist.")
    else:
        if args.dap_path is not None:
            raise Exception("Error: Either --dap or --lengths, not both.")
        for p in args.lengths_paths:
            if not os.path.isfile(p):
                raise Exception("The lengths file %s does not exist." % p)
    if not args.fai_path.endswith(".fai"):
        raise Exception("The fai file has the incorrect file extension.")
    if (args.print_ms + args.print_mems) != 1:
        raise Exception("Error: Either print MSs or MEMs, not both.")
    if args.print_overlaps and args.print_ms:
        raise Exception("Error: Can only print overlaps if printing MEMs.")


def dap_blocks(args):
    """(first position, iterator of int32 [n, C] blocks of consecutive DAP rows): the input is
    streamed like the reference streams it (src/dap_to_bed.py:14-18), never held as a whole."""
    from . import io
    if args.lengths_paths is not None:
        return 0, io.iter_lengths_columns(args.lengths_paths)
    it = io.iter_dap_text(args.dap_path)
    first = next(it, None)
    if first is None:
        return 0, iter(())

    def blocks():
        yield first[1]
        for _, block in it:
            yield block
    return first[0], blocks()


def main(args, sink=None):
    import pyarrow as pa
    from . import api, host, io
    if args.print_ms:
        # dead code in the reference too (NameError at src/dap_to_bed.py:51)
        raise NotImplementedError("--ms is not on the device path (broken in the reference as well)")
    if not args.print_overlaps:
        raise NotImplementedError("only `--mem --overlap` (what `memo index` runs) is on the device path")
    records = api.parse_fai(args.fai_path)
    names = pa.array([r[0] for r in records], type=pa.utf8())
    sink = sys.stdout.buffer if sink is None else sink

    def on_rows(rec_counts, start, end, order):          # BED rows of one chunk, in print order
        io.write_bed(io.index_batch(names, rec_counts, start, end, order), sink)

    pos0, blocks = dap_blocks(args)
    chunk_bytes = int(os.environ.get("MEMO_CHUNK_BYTES", host.DEFAULT_CHUNK_BYTES))
    stats = {}
    host.build_index_streaming(blocks, records, args.sort_lcps, on_rows, pos_first=pos0,
                               chunk_bytes=chunk_bytes, stats=stats)
    if not stats:
        raise KeyError(None)          # reference: fai_dict[None] after an empty DAP


if __name__ == "__main__":
    _args = parse_arguments()
    check_args(_args)
    main(_args)
    sys.stdout.flush()
