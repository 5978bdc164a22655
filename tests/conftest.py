import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    """True if lsw_create succeeds on device 0 (the product has no CPU fallback, so without a device every gpu test would
    fail in lsw_create)."""
    from lamprey_b200 import _native
    _native.load()                       # a missing liblsw.so is NOT a reason to skip: let it raise
    try:
        _native.Engine(0).close()
        return True
    except RuntimeError as e:
        if "no CUDA device" in str(e):
            return False
        raise


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box (pytest -m gpu)")
    for it in gpu_items:
        it.add_marker(skip)


def load_reads(key):
    with gzip.open(os.path.join(GOLD, "reads_%s.txt.gz" % key), "rt") as fp:
        return fp.read().split()


def schema_path(name):
    return os.path.join(GOLD, {"H33": "H3.3_set1.primers", "H31": "H3.1_set1.primers"}[name])


def load_seed_golden(reads_key, schema, thr):
    return np.load(os.path.join(GOLD, "seed_%s_%s_%02d.npz" % (reads_key, schema, round(thr * 100))))


HIT_COLS = ("read", "query", "start", "end", "matches", "mismatches", "score", "q_pos", "q_end", "depth")


def hits_matrix(hits):
    return np.stack([hits[k].astype(np.int32) for k in HIT_COLS], 1) if len(hits) else np.zeros((0, 10), np.int32)


@pytest.fixture(scope="session")
def engine():
    from lamprey_b200 import _native
    return _native.Engine(0)
