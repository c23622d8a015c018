#!/usr/bin/env python3
"""Per-source-line hot spots of an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections, io, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
samples = collections.Counter(); insts = collections.Counter(); src = {}; stalls = collections.defaultdict(collections.Counter)
hdr = None
fname = ""
other = collections.Counter()
main = sys.argv[1:] and "index_build.cu"
for a in sys.argv[3:]:
    if not re.match(r"^\d+-\d+$", a): main = a
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0] in ("File Name", "File Path"):
        fname = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None or len(row) != len(hdr) or not row[0].isdigit(): continue
    if not fname.endswith(main):
        try: other[fname] += int(dict(zip(hdr[4:], row[4:])).get("Instructions Executed") or 0)
        except ValueError: pass
        continue
    ln = int(row[0]); src[ln] = row[1]
    d = dict(zip(hdr[4:], row[4:]))
    try:
        samples[ln] += int(d.get("# Samples") or 0); insts[ln] += int(d.get("Instructions Executed") or 0)
    except ValueError:
        pass
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            try: stalls[ln][k] += int(v)
            except ValueError: pass
tot_s = sum(samples.values()) or 1; tot_i = sum(insts.values()) or 1
print(f"total samples {tot_s}  total warp-insts {tot_i}")
for ln, s in samples.most_common(top):
    st = ",".join(f"{k[6:]}:{v}" for k, v in stalls[ln].most_common(2))
    print(f"{ln:5d} samp {100*s/tot_s:5.1f}%  inst {100*insts[ln]/tot_i:5.1f}%  [{st}]  {src[ln].strip()[:90]}")
# optional: cumulative shares for line ranges "a-b" given as extra args
for f, v in other.most_common(5): print(f"other file {f}: {v} warp-insts")
for rg in sys.argv[3:]:
    if not re.match(r"^\d+-\d+$", rg): continue
    a, b = map(int, rg.split("-"))
    si = sum(v for k, v in insts.items() if a <= k <= b); ss = sum(v for k, v in samples.items() if a <= k <= b)
    print(f"lines {a}-{b}: inst {100*si/tot_i:5.1f}%  samples {100*ss/tot_s:5.1f}%")
