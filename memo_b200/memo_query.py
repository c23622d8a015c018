#!/usr/bin/env python3
"""Drop-in for the reference's src/memo_query.py on the B200 device path.

Same argv (src/memo_query.py:76-87) and the same output file: conservation =
one integer per line, membership (-m) = n space-separated 0/1 per line.

    python -m memo_b200.memo_query [-m] -b index.parquet -r chr:start-end -k K -n N -o out.txt
"""
import argparse


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser(description="k-mer conservation / membership query on a MEMO index "
                                 "(B200 device path).")
    ap.add_argument("-b", "--pq_bed_file", dest="in_file", required=True, help="parquet bed file")
    ap.add_argument("-o", "--out_file", dest="out_file", required=True, help="output file")
    ap.add_argument("-n", "--ndocs", dest="num_docs", required=True,
                    help="total number of genomes in the pangenome")
    ap.add_argument("-k", "--kmer_size", dest="k", required=True, help="k-mer size")
    ap.add_argument("-r", "--genome_region", dest="genome_region", required=True,
                    help="genome region, formatted as chr:start-end")
    ap.add_argument("-m", "--membership_query", dest="membership_query", action="store_true",
                    default=False, help="membership query instead of conservation query")
    return ap.parse_args(argv)


def main(args):
    from . import host, io
    num_docs = int(args.num_docs)
    k = int(args.k)
    record, start_end = args.genome_region.split(":")
    q_start, q_end = map(int, start_end.split("-"))
    f1, f2, f3 = io.read_index_rows(args.in_file, record, q_start, q_end + k)
    text = host.query(f1, f2, f3, q_start, q_end, k, num_docs, args.membership_query, as_text=True)
    with open(args.out_file, "wb") as fh:
        fh.write(text)


if __name__ == "__main__":
    main(parse_arguments())
