/* memo_b200.h -- C ABI of libmemo_b200.so: the B200 (sm_100a) device path for
 * MEMO's DAP -> index conversion and k-mer window query.
 *
 * The reference (StephenHwang/MEMO) has no in-process API: its boundary is
 * argv + files.  The python entry points memo_b200/dap_to_bed.py,
 * parquet_compress_bed.py and memo_query.py keep that CLI boundary; they call the
 * functions below through ctypes.  Each entry point cites the reference lines it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer marked "device" is a CUDA device pointer owned by the caller;
 *    the library allocates nothing persistent
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *    the call returns without synchronising unless stated otherwise
 *  - return value 0 = ok, negative = error; memo_last_error() (thread local)
 *    describes the last failure
 *  - positions are record-relative int32 (chr1 < 2^31); a MEM end p + length is
 *    kept as uint32 (p < 2^31 and length < 2^31 => no overflow)
 *  - DAP values must be >= 0 (matching-statistic lengths); unchecked on device
 */
#ifndef MEMO_B200_H
#define MEMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEMO_B200_ABI_VERSION 8

#define MEMO_OK 0
#define MEMO_ERR_ARG (-1)       /* bad argument */
#define MEMO_ERR_CUDA (-2)      /* CUDA runtime error */
#define MEMO_ERR_UNSUPPORTED (-3)
#define MEMO_ERR_WORKSPACE (-4) /* workspace too small */

/* A run of consecutive DAP rows that belong to one pivot record -- what
 * dap_to_bed.py derives row by row with pos_to_record (src/dap_to_bed.py:76-83)
 * and the header comparison at :121/:125. */
typedef struct memo_segment {
    int64_t row_begin; /* first row of the run in the dap buffer */
    int64_t n_rows;    /* rows in the run (> 0) */
    int32_t pos0;      /* record-relative position of row_begin */
    int32_t rec_len;   /* full record length from the .fai (chr-end row, :126-128) */
    int32_t rec_id;    /* caller's record index; not interpreted */
    int32_t flags;     /* MEMO_SEG_* */
} memo_segment_t;

/* the run starts a record in this buffer: its first row only primes the
 * per-column state (:129-130).  When clear, row_begin >= 1 and row
 * row_begin-1 of the buffer is the halo (the record's previous row, e.g. the
 * last row of the previous position shard). */
#define MEMO_SEG_PRIMED 1
/* emit the chr-end rows after the run's last row (:126-128, :133-134) */
#define MEMO_SEG_CHR_END 2

typedef struct memo_index_opts {
    int32_t order_mode;       /* 1 = --order (conservation), 0 = membership */
    int32_t rows_per_tile;    /* DAP rows per shared-memory stage (general build: per strip), 0 = default */
    int32_t emit_buf_records; /* wide rows: DAP rows per strip (general build: index rows staged per
                                 lane group), 0 = default */
    int32_t warps_per_cta;    /* 0 = default (8) */
    int32_t ctas_per_sm;      /* cap on resident CTAs per SM, 0 = as many as fit */
    int32_t stages;           /* bulk-copy pipeline depth per warp, 0 = default (2), max 4 */
    int32_t kernel_variant;   /* 0 = pick, 1 = strip kernel for narrow rows too (tests), 3 = the
                                 single-kernel strip build (ordered in-place writes through a
                                 look-back over strips; any width) */
    int32_t reserved;
} memo_index_opts_t;

/* slots of the device `result` array written by memo_index_build* */
#define MEMO_RES_N_OUT 0     /* number of index rows the input produces */
#define MEMO_RES_IRREGULAR 1 /* != 0: input is not valid matching statistics;
                                the fast build's output must be discarded and
                                memo_index_build_general run instead */
#define MEMO_RES_REPLAYS 2   /* general build: strips whose staging buffer overflowed (stat) */
#define MEMO_RES_PARKED 3    /* single-kernel strip build: index rows that went through the parking
                                area because an earlier strip was still running (stat) */
#define MEMO_RES_SLOTS 4

int memo_abi_version(void);
const char* memo_last_error(void);

/* Number of SMs of the current device (grid sizing is in multiples of it). */
int memo_device_sm_count(void);

/* Number of kernels this library has launched in the process so far (every
 * launch site counts); reset != 0 returns the count and clears it.  bench.py
 * reports it as gpu_launches. */
int64_t memo_launch_count(int32_t reset);

/* Bytes of device scratch memo_index_build{,_general} need for outputs of
 * out_cap entries (the single-pass build stages its index rows there before
 * the ordered copy into out_*). */
size_t memo_index_workspace_bytes(int64_t rows, int32_t n_cols, int32_t ld, int64_t out_cap,
                                  const memo_segment_t* segs, int32_t n_seg,
                                  const memo_index_opts_t* opts);

/* DAP -> MEMO index rows.  Replaces the hot loop of src/dap_to_bed.py
 * (`--mem --overlap [--order]`): get_new_record :85-91 (row sort), dap_to_mem
 * :116-134 (MEM flag, record transitions, chr-end rows), print_interval /
 * overlaps :93-109 (overlap with the previous MEM of the same column).
 *
 *  dap          device int32 [rows, ld] row-major, 16-byte aligned; the pos
 *               column of dap.txt is NOT stored (rows are consecutive)
 *  segs         HOST array of n_seg runs, in row order, non-overlapping
 *  out_*        device arrays of out_cap entries (start, end, order/genome =
 *               BED f1, f2, f3), written in the reference's print order; may be
 *               NULL with out_cap = 0 to only count
 *  seg_out_end  device int64 [n_seg]: index rows emitted up to and including
 *               run i (so run i owns rows [seg_out_end[i-1], seg_out_end[i]))
 *  result       device int64 [MEMO_RES_SLOTS]
 *
 * Single pass over the DAP (bulk async copies into shared-memory tiles).  Exact
 * for every input for which result[MEMO_RES_IRREGULAR] comes back 0 -- no
 * p + length ever decreases down a column, always the case for matching
 * statistics (MS[p] >= MS[p-1]-1); otherwise the output must be discarded and
 * memo_index_build_general called.  Rows past out_cap are counted but not
 * stored.  n_cols <= 512. */
int memo_index_build(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld,
                     const memo_segment_t* segs, int32_t n_seg,
                     const memo_index_opts_t* opts,
                     int32_t* out_start, uint32_t* out_end, int32_t* out_order,
                     int64_t out_cap, int64_t* seg_out_end, int64_t* result,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Measurement aid: while enabled, every memo_index_build on this host thread
 * brackets its streaming kernel (the launch that reads the DAP) with CUDA events
 * on the build's stream.  memo_profile_collect synchronises those events, returns
 * the summed kernel time and the number of builds, and clears them. */
int memo_profile_enable(int32_t on);
int memo_profile_collect(double* stream_kernel_ms, int32_t* n_builds);

/* Same contract, exact for arbitrary non-negative integer input (three passes:
 * per-strip column aggregates, carry scan, emit). shard_carry_in (device
 * uint32[n_cols] or NULL) is the end of the last flagged MEM per column handed
 * over from the previous position shard for a first run without
 * MEMO_SEG_PRIMED; shard_carry_out (device uint32[n_cols] or NULL) receives the
 * same quantity after the last run (0xFFFFFFFF = no flagged row in this shard). */
int memo_index_build_general(const int32_t* dap, int64_t rows, int32_t n_cols, int32_t ld,
                             const memo_segment_t* segs, int32_t n_seg,
                             const memo_index_opts_t* opts,
                             const uint32_t* shard_carry_in, uint32_t* shard_carry_out,
                             int32_t* out_start, uint32_t* out_end, int32_t* out_order,
                             int64_t out_cap, int64_t* seg_out_end, int64_t* result,
                             void* workspace, size_t workspace_bytes, void* stream);

/* k-mer conservation query over window [q_start, q_end) of one record.
 * Replaces src/memo_query.py memo_init :42-55 (re-centre, shadow cast by k-1,
 * clip, keep end < start), memo_query :57-63 (paint) and the argmax of
 * print_res :70.  Rows are the record's index rows (f1, f2, f3) sorted by f1
 * ascending (index order); the filter of filter_pq :25-27 (q_start < f1 <
 * q_end + k) is applied on device.
 *  out        device uint8 [W] (n_docs <= 255) or uint16 [W] (out_u16 != 0)
 *  status     device int32 [1]: set to 1 if a row has f3 > n_docs (the
 *             reference would write out of bounds)
 *  workspace  memo_query_workspace_bytes(W) bytes */
size_t memo_query_workspace_bytes(int64_t window_len);
int memo_query_conservation(const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                            int64_t n_rows, int64_t q_start, int64_t q_end, int32_t k,
                            int32_t n_docs, void* out, int32_t out_u16, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Membership query (-m): out_bits is device uint32 [W, ceil(n_docs/32)], bit j
 * of a row = column j of the reference's matrix (memo_query.py:51,60-62,68):
 * bit 0 (pivot) set, bit j cleared iff a row with f3 == j covers the k-mer.
 * workspace: memo_query_workspace_bytes(W) bytes (without it the slower tile
 * kernel runs). */
int memo_query_membership(const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                          int64_t n_rows, int64_t q_start, int64_t q_end, int32_t k,
                          int32_t n_docs, uint32_t* out_bits, int32_t* status,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same window answered for several k in ONE launch (BASELINE configs[4]: k = 15 .. 101):
 * tile hand-out, row search and rows (L1/L2) are shared by the k values; src/memo_query.py
 * :42-63 run once per k.  ks: HOST array of n_k (<= 16) values.  Result i (for ks[i]) starts
 * at out + i * stride bytes: conservation (membership == 0, n_docs <= 255) uint8 [W] with
 * stride = W rounded up to 16; membership uint32 [W, ceil(n_docs/32)] with stride = its size
 * (W must then keep that a multiple of 16 bytes). */
int memo_query_sweep(int32_t membership, const int32_t* f1, const uint32_t* f2, const int32_t* f3,
                     int64_t n_rows, int64_t q_start, int64_t q_end, const int32_t* ks, int32_t n_k,
                     int32_t n_docs, void* out, int32_t* status, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Synthetic HPRC-shaped DAP (measurement only): fills rows
 * [row0, row0 + rows) of one record of length rec_len into dap[rows, ld].
 * Integer-only; bit-identical to oracle/memo_oracle.py:synth_dap.
 * workspace: memo_synth_workspace_bytes(rows, n_cols). */
size_t memo_synth_workspace_bytes(int64_t rows, int32_t n_cols);
int memo_synth_dap(int32_t* dap, int64_t row0, int64_t rows, int32_t n_cols, int32_t ld,
                   int64_t rec_len, uint64_t seed, int32_t dense,
                   void* workspace, size_t workspace_bytes, void* stream);

/* BED rows of the index as dap_to_bed.py prints them (src/dap_to_bed.py:100-109,
 * print('\t'.join(map(str, [header, start, end, annot])))): "name TAB f1 TAB f2 TAB f3 LF" per
 * row, for n rows of ONE record (f1 / f2 / f3 device arrays as memo_index_build writes them;
 * `name`: host bytes, at most 256).  out_text (device) needs memo_format_bed_max_bytes(n,
 * name_len) bytes; *out_len (device int64) receives the byte count. */
size_t memo_format_bed_workspace_bytes(int64_t n);
size_t memo_format_bed_max_bytes(int64_t n, int32_t name_len);
int memo_format_bed(const int32_t* f1, const uint32_t* f2, const int32_t* f3, int64_t n, const char* name,
                    int32_t name_len, char* out_text, int64_t* out_len, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Query result text, byte-identical to src/memo_query.py print_res :65-71.
 * conservation: vals device uint8/uint16 [n] -> "%d\n" per value into out_text
 *   (device, capacity >= 4n bytes for uint8, 6n for uint16); *out_len (device
 *   int64) receives the byte count.  workspace: memo_format_workspace_bytes(n).
 * membership: bits device uint32 [W, ceil(n_docs/32)] -> W lines of n_docs
 *   space-separated 0/1 digits; out_text capacity 2 * W * n_docs bytes. */
size_t memo_format_workspace_bytes(int64_t n);
int memo_format_conservation(const void* vals, int32_t is_u16, int64_t n, char* out_text,
                             int64_t* out_len, void* workspace, size_t workspace_bytes,
                             void* stream);
int memo_format_membership(const uint32_t* bits, int64_t W, int32_t n_docs, char* out_text,
                           void* stream);

/* dap.txt text -> DAP rows on the device.  Replaces the reference's line reader and row parse,
 * src/dap_to_bed.py read_file :14-18 and get_new_record :85-88 (`map(int, row.split(' '))`).
 *  text       device bytes [n_bytes] of WHOLE lines "pos v1 .. vC\n" (index.sh:83), 16-byte
 *             aligned, readable up to the next multiple of 32 bytes
 *  pos_first  the position the first line must carry; positions must be consecutive
 *  out        device int32 [max_rows, ld]: row i = the n_cols lengths of line i
 *  result     device int64 [4]: [0] lines parsed, [1] error bits: 1 = a character that is no
 *             digit / space / newline, or an empty field (int() raises in the reference);
 *             2 = a line with another number of fields; 4 = positions not consecutive;
 *             8 = value >= 2^31; 16 = more than max_rows lines
 *  workspace  memo_dap_text_workspace_bytes(n_bytes) */
size_t memo_dap_text_workspace_bytes(int64_t n_bytes);
int memo_dap_text_parse(const uint8_t* text, int64_t n_bytes, int32_t n_cols, int64_t pos_first,
                        int32_t* out, int64_t max_rows, int32_t ld, int64_t* result,
                        void* workspace, size_t workspace_bytes, void* stream);

/* MONI `*.lengths` / `*.lengths.vert` text -> one column of a DAP block, on the HOST (this and
 * memo_lengths_block_parse are the only entry points that do no device work: the tokenizer of the
 * `--lengths` ingest; the block they fill is what goes to the device).
 * Replaces src/index.sh:79 (`grep -v '^>' | tr ' ' '\n' | grep .`), the `paste | nl` of
 * index.sh:83 and the re-parse of src/dap_to_bed.py:87 for one per-genome file.
 *  text         host bytes [n_bytes]: a run of the file, starting where the previous call stopped
 *  final_block  != 0: the file ends with this run (a number touching the end is complete);
 *               0: such a number is left for the next call (result[2] stops before it)
 *  state        in/out, 0 before the first call of a file: 0 = at the start of a line, 1 = inside
 *               a line, 2 = inside a '>' header line (dropped up to its newline)
 *  out          host int32: value i of this call goes to out[i * out_stride]
 *  max_vals     parsing stops before value max_vals + 1
 *  result       host int64 [3]: [0] values written, [1] error bits: 1 = a byte that is no digit,
 *               no white space and no '>' at the start of a line (int() raises in the
 *               reference), 8 = value >= 2^31; [2] bytes consumed
 * One thread per call, no shared state: the host parses the files of a block side by side. */
int memo_lengths_text_parse(const uint8_t* text, int64_t n_bytes, int32_t final_block,
                            int32_t* state, int32_t* out, int64_t out_stride, int64_t max_vals,
                            int64_t* result);

/* A per-genome file of the --lengths ingest as memo_lengths_block_parse walks it: the caller opens
 * the file, owns the buffer (cap bytes) and zeroes everything else before the first call. */
typedef struct {
    int32_t fd;        /* open file descriptor, read with read(2) */
    int32_t state;     /* tokenizer state (memo_lengths_text_parse) */
    int32_t eof;       /* read(2) returned 0 */
    int32_t ended;     /* a tile came back short: the file has no lengths left */
    uint8_t* buf;      /* host buffer */
    int64_t cap;       /* its size; numbers longer than this are errors */
    int64_t lo, hi;    /* unparsed bytes buf[lo, hi) */
    int64_t count;     /* lengths delivered so far */
    int64_t error;     /* 0, or the error bits of memo_lengths_text_parse; 32 = read(2) failed */
} memo_lengths_file_t;

/* One block of the --lengths ingest for a group of neighbouring columns, on the HOST: the next
 * `rows` lengths of files[j] go to out[r * out_stride + j], r = 0 .. rows - 1 (fewer where a file
 * ends: files[j].count tells).  The block is walked in tiles of tile_rows rows, every file of the
 * group filling its column of a tile before the next tile starts, so the tile's cache lines are
 * written while they are resident.  Returns MEMO_ERR_ARG when a file could not be parsed
 * (files[j].error).  One thread per call; calls on disjoint groups run side by side. */
int memo_lengths_block_parse(memo_lengths_file_t* files, int32_t n_files, int32_t* out,
                             int64_t out_stride, int64_t rows, int64_t tile_rows);

/* `memo view` binning.  Replaces src/plot_conservation.py preprocess_data :46-58: per
 * position bin, the number of positions holding each conservation value 0 .. n_docs.
 *  vals    device uint8 [n] (uint16 with is_u16): a conservation vector (memo_query_conservation)
 *  edges   device int64 [n_bins + 1]: bin b = positions [edges[b], edges[b+1]); the host
 *          computes them as the reference does, int(linspace(0, n, n_bins + 1)) (:52)
 *  counts  device uint64 [n_bins, n_docs + 1], zeroed here
 *  status  device int32 [1]: set to 1 if a value above n_docs was seen (not counted) */
int memo_view_bins(const void* vals, int32_t is_u16, int64_t n, int32_t n_docs, int32_t n_bins,
                   const int64_t* edges, uint64_t* counts, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEMO_B200_H */
