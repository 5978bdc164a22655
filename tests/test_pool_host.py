"""CPU tests of the pool's host logic (lamprey_b200/csrc/lsw_pool.cpp) without a GPU: the file is compiled against
tests/pool_stub.cpp, a device-free stand-in for the per-context entry points, and driven through the same C ABI.
Checked: the cut into read ranges and the gather (hits of every range land in global read order), the per-device
turnstiles (uploads and lsw_find_all calls of a device run one at a time and in submission order -- no batch overtakes an
earlier one, nothing deadlocks with many batches in flight), the overflow retry, and the error paths.
Mirrors what /root/reference/lamprey.py:2182-2207 does with processes (deal the reads, collect per worker in order)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from lamprey_b200._native import COUNTERS_DT, HIT16_DT, HIT_DT, POOL_TIMES_DT, _ptr

pytestmark = pytest.mark.timeout(120)

LSW_OK, LSW_E_CUDA, LSW_E_ARG, LSW_E_OVERFLOW, LSW_E_STATE = 0, -1, -2, -4, -5


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pool_stub") / "libpool_stub.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", out,
                           os.path.join(ROOT, "lamprey_b200", "csrc", "lsw_pool.cpp"), os.path.join(ROOT, "tests", "pool_stub.cpp")])
    lib = ctypes.CDLL(out)
    lib.lsw_pool_last_error.restype = ctypes.c_char_p
    lib.stub_log.restype = ctypes.c_int64
    lib.stub_reset.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64]
    return lib


def make_pool(lib, devices, per_dev, pieces=1, compact=False, configure=True):
    h = ctypes.c_void_p(0)
    devs = (ctypes.c_int * len(devices))(*devices)
    assert lib.lsw_pool_create(len(devices), devs, per_dev, ctypes.byref(h)) == LSW_OK
    if configure:
        q = np.frombuffer(b"ACGTACGT", np.uint8)
        qo = np.array([0, 8], np.int32)
        assert lib.lsw_pool_configure(h, 2, -1, -1, -1, 1, _ptr(q), _ptr(qo), ctypes.c_double(0.7), -1, int(compact)) == LSW_OK
        assert lib.lsw_pool_set_pieces(h, pieces) == LSW_OK
    return h


def make_batch(k, n, rng):
    lens = rng.integers(1, 60, n)
    offs = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    text = np.full(int(offs[-1]), ord("A"), np.uint8)
    if len(text):
        text[0] = k                      # the stub logs a batch under the first byte of its text
    return text, offs


def expected_hits(offs):
    """what the stub's contexts produce, in global read order: len % 3 hits per read with start = 0, 1"""
    lens = np.diff(offs)
    reads = np.repeat(np.arange(len(lens)), lens % 3)
    starts = np.concatenate([np.arange(c) for c in lens % 3]) if len(lens) else np.zeros(0, np.int64)
    return reads, starts


def submit(lib, h, text, offs, out):
    t = ctypes.c_int64(0)
    rc = lib.lsw_pool_submit(h, ctypes.c_int64(len(offs) - 1), _ptr(text), _ptr(offs), _ptr(out), ctypes.c_int64(len(out)), ctypes.byref(t))
    return rc, t.value


def wait(lib, h, ticket):
    n = ctypes.c_int64(0)
    cnt = np.zeros(1, COUNTERS_DT)
    tm = np.zeros(1, POOL_TIMES_DT)
    rc = lib.lsw_pool_wait(h, ctypes.c_int64(ticket), ctypes.byref(n), _ptr(cnt), _ptr(tm))
    return rc, n.value, cnt[0], tm[0]


def read_log(lib):
    buf = np.zeros((100000, 4), np.int64)
    n = lib.stub_log(_ptr(buf), ctypes.c_int64(len(buf)))
    return buf[:n]


@pytest.mark.parametrize("devices,per_dev,pieces,compact", [([0], 2, 1, True), ([0, 1, 2], 2, 1, False), ([0, 1, 2], 2, 3, True),
                                                            ([0, 1], 1, 2, False), ([0, 1, 2, 3, 4, 5, 6, 7], 2, 1, True)])
def test_batches_in_flight_are_gathered_in_read_order_and_run_in_submission_order(stub, devices, per_dev, pieces, compact):
    stub.stub_reset(1500, -1, 1 << 30)
    h = make_pool(stub, devices, per_dev, pieces, compact)
    rng = np.random.default_rng(5)
    depth, n_batches = 4, 24
    dt = HIT16_DT if compact else HIT_DT
    batches = [make_batch(k, int(rng.integers(40, 400)), rng) for k in range(n_batches)]
    outs = [np.zeros(1000, dt) for _ in range(depth)]
    tickets = []

    def check(k):
        rc, n, cnt, tm = wait(stub, h, tickets[k])
        assert rc == LSW_OK, stub.lsw_pool_last_error(h)
        reads, starts = expected_hits(batches[k][1])
        got = outs[k % depth][:n]
        assert n == len(reads) and (got["read"] == reads).all() and (got["start"] == starts).all()
        assert cnt["tasks"] == len(batches[k][1]) - 1 and cnt["cells"] == batches[k][1][-1] and cnt["hits"] == n
        assert int(tm["reads"].sum()) == len(batches[k][1]) - 1 and int(tm["hits"].sum()) == n and tm["total_s"] > 0

    for k in range(n_batches):
        if k >= depth:
            check(k - depth)             # the buffer this batch writes into is free again
        rc, t = submit(stub, h, batches[k][0], batches[k][1], outs[k % depth])
        assert rc == LSW_OK
        tickets.append(t)
    for k in range(n_batches - depth, n_batches):
        check(k)
    stub.lsw_pool_destroy(h)

    log = read_log(stub)
    overlaps = (ctypes.c_int * 3)()
    stub.stub_overlaps(overlaps)
    assert list(overlaps) == [0, 0, 0]   # one upload, one lsw_find_all, one download per device at a time
    for d in set(devices):
        for phase in (0, 1):             # uploads and finds of a device: in submission order (batch, then range = read base)
            ev = log[(log[:, 0] == d) & (log[:, 1] == phase)]
            keys = [(int(b), int(base)) for b, base in ev[:, 2:4]]
            assert keys == sorted(keys), (d, phase, keys)
            assert len({k[0] for k in keys}) == n_batches


def test_more_ranges_than_reads_and_an_empty_batch(stub):
    stub.stub_reset(200, -1, 1 << 30)
    h = make_pool(stub, [0, 1, 2], 2, 3)
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 5):
        text, offs = make_batch(n + 1, n, rng)
        out = np.zeros(64, HIT_DT)
        rc, t = submit(stub, h, text, offs, out)
        assert rc == LSW_OK
        rc, nh, cnt, tm = wait(stub, h, t)
        reads, starts = expected_hits(offs)
        assert rc == LSW_OK and nh == len(reads) and (out["read"][:nh] == reads).all() and cnt["tasks"] == n
    stub.lsw_pool_destroy(h)


def test_a_failing_device_fails_the_batch_and_the_pool_stays_usable(stub):
    stub.stub_reset(200, 1, 1 << 30)
    h = make_pool(stub, [0, 1, 2], 2)
    rng = np.random.default_rng(2)
    text, offs = make_batch(1, 300, rng)
    out = np.zeros(1000, HIT_DT)
    tickets = [submit(stub, h, text, offs, out)[1] for _ in range(3)]      # several in flight: none may hang behind the failure
    for t in tickets:
        rc, n, _, _ = wait(stub, h, t)
        assert rc == LSW_E_CUDA and b"device 1" in stub.lsw_pool_last_error(h)
    stub.stub_reset(200, -1, 1 << 30)
    rc, n, _, _ = wait(stub, h, submit(stub, h, text, offs, out)[1])
    assert rc == LSW_OK and n == len(expected_hits(offs)[0])
    assert wait(stub, h, 12345)[0] == LSW_E_ARG                            # unknown ticket
    stub.lsw_pool_destroy(h)


def test_overflow_of_the_caller_buffer_reports_the_needed_count_and_device_overflow_is_retried(stub):
    stub.stub_reset(200, -1, 1 << 30)
    h = make_pool(stub, [0, 1], 1)
    rng = np.random.default_rng(3)
    text, offs = make_batch(1, 500, rng)
    need = len(expected_hits(offs)[0])
    small = np.zeros(need // 2, HIT_DT)
    rc, n, _, _ = wait(stub, h, submit(stub, h, text, offs, small)[1])
    assert rc == LSW_E_OVERFLOW and n == need
    # the DEVICE-side hit buffer too small: lsw_find_all reports the count, the worker makes room and runs it again
    stub.stub_reset(200, -1, 10)
    out = np.zeros(need, HIT_DT)
    rc, n, _, _ = wait(stub, h, submit(stub, h, text, offs, out)[1])
    assert rc == LSW_OK and n == need and (out["read"] == expected_hits(offs)[0]).all()
    stub.lsw_pool_destroy(h)


def test_call_order_and_arguments(stub):
    stub.stub_reset(0, -1, 1 << 30)
    h = make_pool(stub, [0], 1, configure=False)
    text, offs = make_batch(1, 10, np.random.default_rng(4))
    out = np.zeros(100, HIT_DT)
    assert submit(stub, h, text, offs, out)[0] == LSW_E_STATE              # not configured
    q = np.frombuffer(b"ACGT", np.uint8)
    qo = np.array([0, 4], np.int32)
    assert stub.lsw_pool_configure(h, 2, -1, -1, -1, 1, _ptr(q), _ptr(qo), ctypes.c_double(0.0), -1, 0) == LSW_E_ARG     # thr 0 recurses for ever
    assert stub.lsw_pool_configure(h, 2, -1, -1, -1, 1, _ptr(q), _ptr(qo), ctypes.c_double(0.7), -1, 7) == LSW_E_ARG     # unknown record format
    assert stub.lsw_pool_configure(h, 2, -1, -1, -1, 1, _ptr(q), _ptr(qo), ctypes.c_double(0.7), -1, 0) == LSW_OK
    assert stub.lsw_pool_set_pieces(h, 0) == LSW_E_ARG
    stub.stub_reset(20000, -1, 1 << 30)
    rc, t = submit(stub, h, text, offs, out)
    assert rc == LSW_OK
    assert stub.lsw_pool_configure(h, 2, -1, -1, -1, 1, _ptr(q), _ptr(qo), ctypes.c_double(0.7), -1, 0) == LSW_E_STATE   # a batch is in flight
    assert wait(stub, h, t)[0] == LSW_OK
    stub.lsw_pool_destroy(h)
    bad = ctypes.c_void_p(0)
    assert stub.lsw_pool_create(0, (ctypes.c_int * 1)(0), 1, ctypes.byref(bad)) == LSW_E_ARG
    assert stub.lsw_pool_create(1, (ctypes.c_int * 1)(99), 1, ctypes.byref(bad)) == LSW_E_CUDA and not bad.value
