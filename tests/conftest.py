import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def fuzz_golden():
    arrays = np.load(os.path.join(GOLDEN, "fuzz.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "fuzz.json")))
    return arrays, meta


@pytest.fixture(scope="session")
def example_golden():
    rd = lambda n: open(os.path.join(GOLDEN, n)).read()
    dap_txt = rd("example.dap.txt")
    rows = np.array([[int(x) for x in l.split()] for l in dap_txt.splitlines()], dtype=np.int64)
    fai = [(l.split()[0], int(l.split()[1])) for l in rd("example.fai").splitlines()]
    return {
        "dap_txt": dap_txt, "pos": rows[:, 0], "vals": rows[:, 1:], "records": fai,
        "cons_bed": rd("example.cons.bed"), "memb_bed": rd("example.memb.bed"),
        "queries": json.load(open(os.path.join(GOLDEN, "example.queries.json"))),
    }
