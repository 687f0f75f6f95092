"""CPU test of the host half of the byte transfer of whole-number float batches (csrc/mlp.cu: step_packed / bla_pack_pixels): the
packing pool's output layout, and that anything that is not bit for bit a whole number 0..255 marks its chunk -- and only its chunk --
as inexact (the library then sends that chunk as floats).  Pixels arrive as such floats from the reference's CSV loader
(lib/mnist_csv2.c:13-34, model/mnist_nn.c:204-209)."""
import ctypes as C
import time

import numpy as np
import pytest

from helpers import ptr


@pytest.fixture(scope="module")
def bla():
    import bla_b200 as b
    return b


def pack(b, x, chunk):
    rows, cols = x.shape
    chunks = -(-cols // chunk)
    out = np.full(rows * cols, 0xAB, np.uint8)
    exact = np.full(chunks, -1, np.int32)
    r = b.bla_pack_pixels(ptr(x), rows, cols, chunk, ptr(out), ptr(exact))
    return r, out, exact


def expected(x, chunk):
    rows, cols = x.shape
    parts = [np.ascontiguousarray(x[:, b0:b0 + chunk]).astype(np.uint8).ravel() for b0 in range(0, cols, chunk)]
    return np.concatenate(parts)


def test_whole_number_batches_pack_into_chunk_contiguous_bytes(bla):
    rng = np.random.default_rng(5)
    for rows, cols, chunk in ((784, 5000, 1024), (784, 4096, 4096), (17, 130, 64), (1, 1, 64)):
        x = rng.integers(0, 256, (rows, cols)).astype(np.float32)
        r, out, exact = pack(bla, x, chunk)
        assert r == 1 and exact.tolist() == [1] * len(exact)
        assert np.array_equal(out, expected(x, chunk)), (rows, cols, chunk)


@pytest.mark.parametrize("poison", [0.5, -0.0, 256.0, -1.0, 255.000015, np.nan, np.inf, 1e20, -3e9])
def test_any_other_value_marks_its_chunk_only(bla, poison):
    rng = np.random.default_rng(6)
    x = rng.integers(0, 256, (784, 5000)).astype(np.float32)
    x[400, 2500] = poison                      # chunk 2 of 1024-column chunks
    r, out, exact = pack(bla, x, 1024)
    assert r == 0 and exact.tolist() == [1, 1, 0, 1, 1]
    want = expected(np.where(np.isfinite(x) & (x >= 0) & (x < 256), x, 0).astype(np.float32), 1024)
    sel = np.ones(out.size, bool)
    sel[784 * 2048:784 * 3072] = False         # the inexact chunk's bytes mean nothing
    assert np.array_equal(out[sel], want[sel])


def test_pool_is_reusable_and_not_slower_than_numpy(bla):
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, (784, 12288)).astype(np.float32)
    for _ in range(3):                         # the same persistent pool, job after job
        r, out, exact = pack(bla, x, 6144)
        assert r == 1
    t0 = time.perf_counter(); pack(bla, x, 6144); t_pool = time.perf_counter() - t0
    t0 = time.perf_counter(); x.astype(np.uint8); t_np = time.perf_counter() - t0
    print(f"pack 784 x 12288: pool {t_pool * 1e3:.2f} ms, numpy astype {t_np * 1e3:.2f} ms")
