"""Multi-GPU parity under pytest: spawns one rank per visible GPU (torch.distributed.run,
NCCL) running scripts/mgpu_check.py -- sharded index build incl. the irregular carry
exchange, gathered in rank order and compared with the oracle's unsharded index.
Skipped on boxes with fewer than two GPUs (the CPU/gloo twin is test_sharded_gloo.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_build_nccl_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "scripts", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "mgpu_check PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_dap_to_bed_gpus_n_writes_the_same_bytes(tmp_path):
    """The multi-GPU product entry: `dap_to_bed --gpus N` (position shards = byte shares of
    dap.txt, one rank per GPU, ordered write) produces the bytes of the one-GPU run and of the
    oracle; input that is not matching statistics falls back to the exact single-rank build."""
    import numpy as np
    import torch
    from oracle import memo_oracle as mo
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least two GPUs")
    lens = [60000, 1, 45000, 30000]
    recs = [(f"c{i}", n) for i, n in enumerate(lens)]
    fai = tmp_path / "p.fa.fai"
    fai.write_text("".join(f"{h}\t{n}\t7\t{n}\t{n + 1}\n" for h, n in recs))
    rng = np.random.default_rng(2)
    valid = np.concatenate([mo.synth_dap(n, 9, seed=70 + i, dense=(i == 2)) for i, n in enumerate(lens)])[:-321]
    junk = valid.copy()
    junk[70000:70500] = rng.integers(0, 50, (500, 9))
    env = dict(os.environ, MEMO_CHUNK_BYTES=str(1 << 20), PYTHONPATH=ROOT)
    for name, vals in (("valid", valid), ("irregular", junk)):
        dap = tmp_path / f"{name}.txt"
        dap.write_text("".join(f"{i} " + " ".join(map(str, row)) + "\n" for i, row in enumerate(vals)))
        for order in (True, False):
            want = mo.format_bed(recs, *mo.index_build(vals, recs, order))
            for n in sorted({1, 2, n_gpu}):
                out = tmp_path / f"{name}_{order}_{n}.bed"
                cmd = [sys.executable, "-m", "memo_b200.dap_to_bed", "--mem", "--overlap", "--fai", str(fai),
                       "--dap", str(dap), "--gpus", str(n), "--out", str(out)] + (["--order"] if order else [])
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
                assert r.returncode == 0, r.stderr[-3000:]
                assert out.read_text() == want, (name, order, n)
    # without --out the payload still goes to stdout (index.sh redirects it)
    cmd = [sys.executable, "-m", "memo_b200.dap_to_bed", "--mem", "--overlap", "--order", "--fai", str(fai),
           "--dap", str(tmp_path / "valid.txt"), "--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout == mo.format_bed(recs, *mo.index_build(valid, recs, True)), r.stderr[-2000:]
