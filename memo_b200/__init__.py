"""memo_b200 -- B200-native (sm_100a) device path for MEMO's hot path:
DAP -> conservation/membership index rows, and k-mer window queries.

Drop-in entry points with the reference's argv:
  python -m memo_b200.dap_to_bed            (src/dap_to_bed.py)
  python -m memo_b200.parquet_compress_bed  (src/parquet_compress_bed.py)
  python -m memo_b200.memo_query            (src/memo_query.py)
"""
__version__ = "0.1.0"
