#!/usr/bin/env python3
"""Drop-in for the reference's src/dap_to_bed.py on the B200 device path.

Same argv (src/dap_to_bed.py:139-149), same validation (:151-171), BED payload
on stdout and nothing else (index.sh:93,102 redirect stdout into the .bed).

    python -m memo_b200.dap_to_bed --mem [--order] --overlap --fai P.fai --dap dap.txt > out.bed

Extensions (not in the reference):
  --lengths G2.lengths G3.lengths ...   instead of --dap: reads the per-genome MONI outputs
      directly (genome_list.txt order minus the pivot) and skips index.sh:79-83's vertical
      files, `paste | nl` and the text re-parse;
  --gpus N [--out P.bed]   the pivot is cut into N position shards (equal byte shares of
      dap.txt), one process per GPU (torch.distributed.run, NCCL); every rank builds the rows
      of its shard from its own lines plus the line before them, the byte counts of the BED
      parts are all-gathered and every rank writes its part at its offset: the same bytes as
      one GPU writes.  Without --out the result still goes to stdout;
  --parquet P.parquet   the index rows go from the device columns straight into the Parquet
      index (schema, codec and row order of parquet_compress_bed.py; row groups cut at record
      changes with min/max statistics), no BED text in between.
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
    ap.add_argument("--gpus", dest="gpus", type=int, default=1,
                    help="extension: position-shard the DAP over N GPUs of this box (needs --dap)")
    ap.add_argument("--out", dest="out_path", default=None,
                    help="extension: write the BED to this file instead of stdout")
    ap.add_argument("--parquet", dest="parquet_path", default=None,
                    help="extension: write the Parquet index directly (no BED text, no parquet_compress_bed)")
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
    if args.gpus > 1 and args.dap_path is None:
        raise Exception("Error: --gpus needs --dap (position shards are byte shares of dap.txt).")
    if args.parquet_path is not None and (args.gpus > 1 or args.out_path is not None):
        raise Exception("Error: --parquet is a single-process output of its own (no --gpus, no --out).")


def dap_blocks(args, byte_range=None):
    """(first position, iterator of int32 [n, C] blocks of consecutive DAP rows): the input is
    streamed like the reference streams it (src/dap_to_bed.py:14-18), never held as a whole.
    dap.txt is parsed on the device (the host only moves bytes; MEMO_TEXT_PARSE=host selects the
    pyarrow parser); the .lengths files are parsed on the host."""
    from . import io
    if args.lengths_paths is not None:
        return 0, io.iter_lengths_columns(args.lengths_paths)
    if os.environ.get("MEMO_TEXT_PARSE", "device") == "host":
        it = io.iter_dap_text(args.dap_path, byte_range=byte_range)
    else:
        it = io.iter_dap_text_device(args.dap_path, byte_range=byte_range,
                                     block_bytes=int(os.environ.get("MEMO_TEXT_BLOCK_BYTES", 128 << 20)))
    first = next(it, None)
    if first is None:
        return 0, iter(())

    def blocks():
        yield first[1]
        for _, block in it:
            yield block
    return first[0], blocks()


def _stream(args, records, sink, pos0, blocks, **kw):
    """DAP blocks -> BED rows on `sink`, chunk by chunk.  Returns the stream's stats."""
    import pyarrow as pa
    from . import api, host, io
    to_parquet = isinstance(sink, io.IndexParquetWriter)
    # BED text is formatted on the device (MEMO_BED_FORMAT=host: the Arrow CSV writer)
    on_device = not to_parquet and os.environ.get("MEMO_BED_FORMAT", "device") != "host" and \
        all(len(r[0].encode("utf-8")) <= 256 for r in records)
    names = pa.array([r[0] for r in records], type=pa.utf8())
    fmt = api.BedFormatter() if on_device else None

    def on_rows_device(rec_counts, dev_rows):            # rows of one chunk on the device, in print order
        a = 0
        for rid, cnt in rec_counts:
            sink.write(memoryview(fmt.format(dev_rows[:, a:a + cnt], records[rid][0])))
            a += cnt

    def on_rows_host(rec_counts, start, end, order):     # rows of one chunk, in print order
        table = io.index_batch(names, rec_counts, start, end, order)
        if to_parquet:
            sink.write(table)
        else:
            io.write_bed(table, sink)

    on_rows = on_rows_device if on_device else on_rows_host
    kw = dict(kw, on_device=on_device)
    chunk_bytes = int(os.environ.get("MEMO_CHUNK_BYTES", host.DEFAULT_CHUNK_BYTES))
    stats = {}
    host.build_index_streaming(blocks, records, args.sort_lcps, on_rows, pos_first=pos0,
                               chunk_bytes=chunk_bytes, stats=stats, **kw)
    return stats


def main(args, sink=None):
    from . import api
    if args.print_ms:
        # dead code in the reference too (NameError at src/dap_to_bed.py:51)
        raise NotImplementedError("--ms is not on the device path (broken in the reference as well)")
    if not args.print_overlaps:
        raise NotImplementedError("only `--mem --overlap` (what `memo index` runs) is on the device path")
    records = api.parse_fai(args.fai_path)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and "LOCAL_RANK" in os.environ:
        return main_rank(args, records)                   # one of the N processes of --gpus N
    if args.gpus > 1:
        return main_spawn(args, sink)
    from . import io
    out = None
    if args.parquet_path is not None:
        # index rows straight from the device columns into Parquet (skips the BED text and its re-parse)
        out = io.IndexParquetWriter(args.parquet_path)
    elif args.out_path and sink is None:
        out = open(args.out_path, "wb")
    try:
        pos0, blocks = dap_blocks(args)
        stats = _stream(args, records, out or sink or sys.stdout.buffer, pos0, blocks)
    finally:
        if out is not None:
            out.close()
    if not stats:
        raise KeyError(None)          # reference: fai_dict[None] after an empty DAP


def main_spawn(args, sink=None):
    """--gpus N outside torchrun: start the N ranks, then pass the result on."""
    import shutil
    import socket
    import subprocess
    import tempfile
    out_path = args.out_path
    tmp = None
    if out_path is None:
        fd, tmp = tempfile.mkstemp(suffix=".bed", dir=os.path.dirname(os.path.abspath(args.dap_path)))
        os.close(fd)
        out_path = tmp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    argv = ["--fai", args.fai_path, "--dap", args.dap_path, "--mem", "--overlap", "--gpus", str(args.gpus),
            "--out", out_path] + (["--order"] if args.sort_lcps else [])
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "memo_b200.dap_to_bed"] + argv
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:
            sys.stderr.buffer.write(r.stderr[-8000:])
            raise Exception(f"sharded index build failed (exit {r.returncode})")
        if tmp is not None:
            with open(tmp, "rb") as fh:
                shutil.copyfileobj(fh, sink or sys.stdout.buffer, 16 << 20)
    finally:
        if tmp is not None and os.path.exists(tmp):
            os.remove(tmp)


def main_rank(args, records):
    """One rank of a sharded build: its byte share of dap.txt (+ the line before it), BED rows
    into a part file, all-gather of the part sizes, ordered write into --out."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import io, shard
    if args.out_path is None:
        raise Exception("Error: a sharded build needs --out.")
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        lo, hi, prev = io.split_text_rows(args.dap_path, world, rank)
        pos0, blocks = dap_blocks(args, byte_range=(lo, hi))
        part = f"{args.out_path}.part{rank:03d}"
        with open(part, "wb") as fh:
            halo = None if prev is None else np.array([int(x) for x in prev.split(b" ")][1:], dtype=np.int32)
            stats = _stream(args, records, fh, pos0, blocks, first_halo=halo, final=(rank == world - 1))
        # input that is not matching statistics needs the carry of everything before a shard:
        # one rank redoes the whole file with the exact streaming build
        flag = torch.tensor([1 if stats.get("general") else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            os.remove(part)
            if rank == 0:
                pos0, blocks_all = dap_blocks(args)
                with open(args.out_path, "wb") as fh:
                    _stream(args, records, fh, pos0, blocks_all)
            dist.barrier()
        else:
            shard.ordered_file_write(part, args.out_path, dev)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    _args = parse_arguments()
    check_args(_args)
    main(_args)
    sys.stdout.flush()
