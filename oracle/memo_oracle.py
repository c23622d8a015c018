"""CPU oracle for the MEMO hot path (TEST INFRASTRUCTURE -- not product code).

This is a from-scratch numpy restatement of the algorithm the reference runs
in three scripts.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``memo_b200``) never does.

Parity status: PINNED.  The restatement is checked against outputs of the
reference scripts themselves (``tests/golden/*.npz``, produced by
``tests/golden/make_goldens.py`` running ``/root/reference/src/*.py``
unmodified in the build container), including the SURVEY A.4 example vectors.

Reference lines each function follows (paths relative to the reference root):

* ``parse_fai``            -> ``src/dap_to_bed.py:20-28``
* ``split_runs``           -> ``src/dap_to_bed.py:76-91,119-131`` (position ->
  record lookup, record transitions)
* ``index_build``          -> ``src/dap_to_bed.py:85-134`` (``--mem --overlap``
  with/without ``--order``)
* ``index_build_stream``   -> same lines, literal row-at-a-time form (small
  cases only; cross-checks the closed form)
* ``query_select``         -> ``src/memo_query.py:19-36`` (parquet predicates)
* ``query``                -> ``src/memo_query.py:42-71`` (shadow cast, clip,
  paint, argmax / matrix)
* ``format_*``             -> ``src/dap_to_bed.py:106``, ``src/memo_query.py:65-71``

Closed form used by ``index_build`` (one record run, rows r = 0..n-1 at
record-relative positions p = p0 + r, C columns):

  S[r]   = sort_desc(v[r]) if order else v[r]
  E[r,j] = p + S[r,j]                       ("MEM end" of cell (r, j))
  flag   = r == 0  or  E[r,j] > E[r-1,j]    (== S[r-1,j] <= S[r,j])
  c[r,j] = E at the last flagged row <= r   (the dict entry of print_interval)
  emit (p, min(c[r-1,j], E[r,j]), j+1) for flagged r > 0 iff that min >= p
  after the last row: (n_rec, min(c[n-1,j], 2*n_rec), j+1) iff >= n_rec
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "parse_fai", "split_runs", "index_build", "index_build_stream",
    "query_select", "query", "format_bed", "format_conservation",
    "format_membership", "synth_dap", "GRCH38_LENGTHS",
]


# --------------------------------------------------------------------------
# .fai and record runs
# --------------------------------------------------------------------------
def parse_fai(path):
    """[(header, length)] from the first two whitespace columns of a .fai
    (``dap_to_bed.py:20-28``)."""
    out = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            header, length, *_ = line.split()
            out.append((header, int(length)))
    return out


def split_runs(pos, records):
    """Split the DAP rows into maximal runs of consecutive rows that map to
    the same record (``dap_to_bed.py:76-83,121,125``).

    pos      int64[L] global positions (first DAP column)
    records  [(header, length)]
    returns  [(rec_idx, row_begin, row_end)] and int64[L] record-relative pos
    Raises like the reference when a position lies beyond every record.
    """
    pos = np.asarray(pos, dtype=np.int64)
    lens = np.array([r[1] for r in records], dtype=np.int64)
    ends = np.cumsum(lens)
    starts = ends - lens
    if pos.size == 0:
        return [], pos
    rec = np.searchsorted(ends, pos, side="right")
    if (rec >= len(records)).any() or (pos < 0).any():
        raise Exception("Position beyond all intervals; ensure your fai file is "
                        "from fasta of initial query.")
    # zero-length records can never own a position; searchsorted(side=right)
    # already skips them because start <= pos < end is empty for them.
    rel = pos - starts[rec]
    cut = np.flatnonzero(np.diff(rec) != 0) + 1
    bounds = np.concatenate([[0], cut, [pos.size]])
    runs = [(int(rec[bounds[i]]), int(bounds[i]), int(bounds[i + 1]))
            for i in range(len(bounds) - 1)]
    return runs, rel


# --------------------------------------------------------------------------
# index build  (dap_to_bed.py --mem --overlap [--order])
# --------------------------------------------------------------------------
def _run_build(vals, rel, rec_len, order):
    """One record run. vals int64[n, C]; rel int64[n] record-relative positions.
    Returns (start, end, col) int64 arrays in reference print order."""
    n, C = vals.shape
    S = -np.sort(-vals, axis=1) if order else vals
    E = rel[:, None] + S
    flag = np.ones((n, C), dtype=bool)
    flag[1:] = S[:-1] <= S[1:]                      # dap_to_bed.py:123
    rows = np.arange(n, dtype=np.int64)[:, None]
    last = np.maximum.accumulate(np.where(flag, rows, -1), axis=0)
    cols = np.arange(C)[None, :]
    c_end = E[last, cols]                           # end of the stored MEM
    c_start = rel[last]                             # its start
    emit = np.zeros((n, C), dtype=bool)
    end = np.zeros((n, C), dtype=np.int64)
    start = np.zeros((n, C), dtype=np.int64)
    if n > 1:
        start[1:] = np.maximum(c_start[:-1], rel[1:, None])   # :95
        end[1:] = np.minimum(c_end[:-1], E[1:])               # :96
        emit[1:] = flag[1:] & (end[1:] >= start[1:])          # :97
    r_idx, c_idx = np.nonzero(emit)                 # row-major == print order
    o_start = start[r_idx, c_idx]
    o_end = end[r_idx, c_idx]
    o_col = c_idx.astype(np.int64) + 1
    # chr-end sentinels (:126-128, :133-134)
    s_start = np.maximum(c_start[-1], rec_len)
    s_end = np.minimum(c_end[-1], 2 * rec_len)
    keep = s_end >= s_start
    o_start = np.concatenate([o_start, s_start[keep]])
    o_end = np.concatenate([o_end, s_end[keep]])
    o_col = np.concatenate([o_col, (np.arange(C, dtype=np.int64) + 1)[keep]])
    return o_start, o_end, o_col


def index_build(dap, records, order, pos=None):
    """MEMO index rows from a DAP.

    dap      int[L, C] matching-statistic lengths (DAP without the pos column)
    records  [(header, length)] in .fai order
    order    True = conservation index (--order), False = membership index
    pos      optional int[L] first DAP column; default 0..L-1 (index.sh:83)
    returns  (rec_idx, start, end, col) int64 arrays, reference print order
    """
    dap = np.asarray(dap, dtype=np.int64)
    L = dap.shape[0]
    if pos is None:
        pos = np.arange(L, dtype=np.int64)
    runs, rel = split_runs(pos, records)
    out_r, out_s, out_e, out_c = [], [], [], []
    for rec, b, e in runs:
        s_, e_, c_ = _run_build(dap[b:e], rel[b:e], records[rec][1], order)
        out_r.append(np.full(s_.shape, rec, dtype=np.int64))
        out_s.append(s_); out_e.append(e_); out_c.append(c_)
    if not out_r:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), z.copy(), z.copy()
    return (np.concatenate(out_r), np.concatenate(out_s),
            np.concatenate(out_e), np.concatenate(out_c))


def index_build_stream(dap, records, order, pos=None):
    """Literal row-at-a-time form of the same lines (small inputs only)."""
    dap = np.asarray(dap, dtype=np.int64)
    L = dap.shape[0]
    if pos is None:
        pos = np.arange(L, dtype=np.int64)
    bounds = []
    acc = 0
    for h, n in records:
        bounds.append((acc, acc + n)); acc += n
    rows = []
    prev_rec, prev_row, stored = None, None, {}

    def put(rec, start, end, col):
        old = stored.get(col)
        if old is not None:
            s, e = max(old[0], start), min(old[1], end)
            if e >= s:
                rows.append((rec, s, e, col))
        stored[col] = (start, end)

    C = dap.shape[1] if L else 0
    for r in range(L):
        g = int(pos[r])
        rec = None
        for i, (a, b) in enumerate(bounds):
            if a <= g < b:
                rec = i; break
        if rec is None:
            raise Exception("Position beyond all intervals; ensure your fai file "
                            "is from fasta of initial query.")
        p = g - bounds[rec][0]
        cur = sorted((int(x) for x in dap[r]), reverse=True) if order \
            else [int(x) for x in dap[r]]
        if rec == prev_rec:
            for j in range(C):
                if prev_row[j] <= cur[j]:
                    put(rec, p, p + cur[j], j + 1)
        else:
            if prev_rec is not None:
                n = records[prev_rec][1]
                for j in range(C):
                    put(prev_rec, n, 2 * n, j + 1)
            stored.clear()
            for j in range(C):
                put(rec, p, p + cur[j], j + 1)
        prev_rec, prev_row = rec, cur
    if prev_rec is not None:
        n = records[prev_rec][1]
        for j in range(C):
            put(prev_rec, n, 2 * n, j + 1)
    a = np.array(rows, dtype=np.int64).reshape(-1, 4)
    return a[:, 0], a[:, 1], a[:, 2], a[:, 3]


# --------------------------------------------------------------------------
# query  (memo_query.py)
# --------------------------------------------------------------------------
def query_select(f1, f2, f3, q_start, q_end_plus_k):
    """Rows the two parquet predicates return, concatenated in the reference's
    order: [f1 <= s < f2] then [s < f1 < e + k]  (``memo_query.py:22-36``)."""
    f1 = np.asarray(f1, dtype=np.int64)
    f2 = np.asarray(f2, dtype=np.int64)
    f3 = np.asarray(f3, dtype=np.int64)
    a = (f1 <= q_start) & (f2 > q_start)
    b = (f1 > q_start) & (f1 < q_end_plus_k)
    sel = np.concatenate([np.flatnonzero(a), np.flatnonzero(b)])
    return f1[sel], f2[sel], f3[sel]


def query(f1, f2, f3, q_start, q_end, k, n_docs, membership):
    """k-mer conservation / membership over [q_start, q_end) of one record.

    f1, f2, f3  the index rows of that record (start, end, order/genome)
    returns     conservation: int64[W]; membership: uint8[W, n_docs]
    """
    W = q_end - q_start
    s1, s2, s3 = query_select(f1, f2, f3, q_start, q_end + k)   # :100
    start = np.clip(s1 - q_start, 0, W)                         # :46,48
    cend = np.clip(s2 - q_start - (k - 1), 0, W)                # :46-48
    keep = cend < start                                         # :49
    start, cend, col = start[keep], cend[keep], s3[keep]
    width = n_docs if membership else n_docs + 1
    if col.size and (col.max() >= width or col.min() < 0):
        raise IndexError("index row order/genome id out of range for -n")
    if membership:
        rec = np.ones((W, n_docs), dtype=bool)                  # :51
        bit = False
    else:
        rec = np.zeros((W, n_docs + 1), dtype=bool)             # :53-54
        rec[:, n_docs] = True
        bit = True
    for a, b, j in zip(cend.tolist(), start.tolist(), col.tolist()):
        rec[a:b, j] = bit                                       # :61-62
    if membership:
        return rec.astype(np.uint8)                             # :68
    return np.argmax(rec, axis=1).astype(np.int64)              # :70


# --------------------------------------------------------------------------
# text formats
# --------------------------------------------------------------------------
def format_bed(records, rec_idx, start, end, col):
    """BED payload exactly as ``dap_to_bed.py:106`` prints it."""
    names = [r[0] for r in records]
    return "".join(f"{names[r]}\t{s}\t{e}\t{c}\n"
                   for r, s, e, c in zip(rec_idx.tolist(), start.tolist(),
                                         end.tolist(), col.tolist()))


def format_conservation(vec):
    """``print(*rec, sep='\\n', file=...)`` (``memo_query.py:71``)."""
    return "\n".join(str(int(x)) for x in vec) + "\n"


def format_membership(mat):
    """``np.savetxt(..., delimiter=' ', fmt='%i')`` (``memo_query.py:68``)."""
    return "".join(" ".join(str(int(x)) for x in row) + "\n" for row in mat)


# --------------------------------------------------------------------------
# `memo view` binning (src/plot_conservation.py:46-58)
# --------------------------------------------------------------------------
def view_bins(vec, n_docs, n_bins):
    """float64 [n_bins, n_docs + 1]: per bin, count(value) / bin size, bins cut at
    int(linspace(0, positions, n_bins + 1)) (``plot_conservation.py:51-56``).  Raises
    ZeroDivisionError on an empty bin like the reference."""
    vec = [int(x) for x in vec]
    positions = len(vec)
    edges = list(map(int, np.linspace(0, positions, n_bins + 1)))
    out = np.zeros((n_bins, n_docs + 1), dtype=np.float64)
    for b, (lo, hi) in enumerate(zip(edges[:-1], edges[1:])):
        chunk = vec[lo:hi]
        total = len(chunk)
        for order in range(n_docs + 1):
            out[b, order] = chunk.count(order) / total
    return out


# --------------------------------------------------------------------------
# synthetic HPRC-shaped DAP.  Integer-only so that this numpy
# form and the CUDA generator agree bit for bit.
# --------------------------------------------------------------------------
GRCH38_LENGTHS = [
    248956422, 242193529, 198295559, 190214555, 181538259, 170805979,
    159345973, 145138636, 138394717, 133797422, 135086622, 133275309,
    114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
    58617616, 64444167, 46709983, 50818468, 156040895, 57227415,
]

_K1 = np.uint64(0x9E3779B97F4A7C15)
_K2 = np.uint64(0xC2B2AE3D27D4EB4F)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _ctz(x):
    # count trailing zeros of non-zero uint64 values
    x = x.astype(np.uint64)
    low = x & (~x + np.uint64(1))
    # exact for powers of two below 2**53
    return np.log2(low.astype(np.float64)).astype(np.int64)


def synth_draws(p, C, seed, dense=False):
    """d[p, c]: the i.i.d. integer draws.  p int64[n] record-relative rows."""
    with np.errstate(over="ignore"):
        pp = p.astype(np.uint64)[:, None] * _K1
        cc = np.arange(C, dtype=np.uint64)[None, :] * _K2
        h = _splitmix64(np.uint64(seed) ^ pp ^ cc)
    d_s = 12 + _ctz(h | np.uint64(1 << 20))
    if dense:
        is_long = ((h >> np.uint64(20)) & np.uint64(1023)) == 0
        d_l = ((_ctz((h >> np.uint64(32)) | np.uint64(1 << 16)) + 1) << 8) + \
            ((h >> np.uint64(48)) & np.uint64(255)).astype(np.int64)
    else:
        is_long = ((h >> np.uint64(20)) & np.uint64(511)) == 0
        d_l = ((_ctz((h >> np.uint64(32)) | np.uint64(1 << 16)) + 1) << 10) + \
            ((h >> np.uint64(48)) & np.uint64(1023)).astype(np.int64)
    return np.where(is_long, np.maximum(d_s, d_l), d_s).astype(np.int64)


def synth_dap(rec_len, C, seed, row0=0, rows=None, dense=False):
    """Valid matching statistics for rows [row0, row0+rows) of one record:
    MS[p, c] = min(max_{q<=p}(d[q, c] + q) - p, rec_len - p).
    The draw is bounded (< 2**15), so the prefix maximum only needs a look-back
    window of LOOKBACK rows; rows before row0 - LOOKBACK cannot matter."""
    LOOKBACK = 32768
    if rows is None:
        rows = rec_len - row0
    lo = max(0, row0 - LOOKBACK)
    p = np.arange(lo, row0 + rows, dtype=np.int64)
    d = synth_draws(p, C, seed, dense)
    reach = np.maximum.accumulate(d + p[:, None], axis=0)
    ms = np.minimum(reach - p[:, None], rec_len - p[:, None])
    return ms[row0 - lo:].astype(np.int32)
