"""GPU parity of the device data pipeline (csrc/data.cu) with lib/mnist_csv2.c, lib/cifar10.c and the epoch loop
of model/mnist_nn.c:181-342: same sample order under the same libc rand() stream, same batch matrices, same loss curve as the
reference program started from the same `init` checkpoint (which also exercises the CSV codec, N2)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import REF_DIR, ptr
from test_data_cpu import libc, reference_take_sequence
import test_programs_gpu as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bla():
    import bla_b200 as b
    assert b.bla_device_count() >= 1
    b.bla_set_gemm_path(b.GEMM_FP32)
    return b


def ref_data():
    p = os.path.join(REF_DIR, "libref_data.so")
    return C.CDLL(p) if os.path.exists(p) else None


class MnistCSV(C.Structure):        # lib/mnist_csv2.h:5-12
    _fields_ = [("file", C.c_void_p), ("X", C.POINTER(C.c_float)), ("y", C.POINTER(C.c_float)), ("num_examples", C.c_int),
                ("num_sampled", C.c_int), ("sampled", C.c_void_p)]


class MnistExample(C.Structure):    # lib/mnist_csv2.h:14-18
    _fields_ = [("X", C.POINTER(C.c_float)), ("y", C.c_float), ("num_examples", C.c_int)]


@pytest.mark.parametrize("n,draws", [(50, 50), (1000, 2500), (7, 30)])
def test_sampler_order_matches_the_reference(bla, n, draws):
    b = bla
    x = np.zeros((n, 4), np.float32); y = np.zeros(n, np.float32)
    store = b.bla_mnist_from_arrays(ptr(x), ptr(y), n, 4)
    got = np.empty(draws, np.int32)
    libc.srand(42)
    b.bla_mnist_sample_take(store, draws, ptr(got))
    assert got.tolist() == reference_take_sequence(n, draws, 42)
    ref = ref_data()
    if ref is not None:                                  # the compiled lib/mnist_csv2.c itself
        ref.get_random_data_take.restype = MnistExample
        ref.get_random_data_take.argtypes = [C.POINTER(MnistCSV)]
        X = np.zeros(n * 4, np.float32); Y = np.zeros(n, np.float32); flags = np.zeros(n + 8, np.uint8)
        csv = MnistCSV(None, X.ctypes.data_as(C.POINTER(C.c_float)), Y.ctypes.data_as(C.POINTER(C.c_float)), n, 0, flags.ctypes.data)
        libc.srand(42)
        want = []
        for _ in range(draws):
            ex = ref.get_random_data_take(C.byref(csv))
            want.append((C.addressof(ex.X.contents) - X.ctypes.data) // 4)
        assert got.tolist() == want
    b.bla_mnist_destroy(store)


def test_csv_store_and_gather_match_the_batch_assembly(bla, tmp_path):
    b = bla
    path = tmp_path / "mnist.csv"
    P.mnist_csv(path, 300, 5)
    rows = np.array([[float(v) for v in line.rstrip(",\n").split(",")] for line in open(path)], np.float32)
    store = b.bla_mnist_from_csv(str(path).encode())
    assert b.bla_mnist_num_examples(store) == 300
    idx = np.random.default_rng(1).integers(0, 300, 77).astype(np.int32)
    xd = b.bla_malloc_device(784 * 77 * 4); yd = b.bla_malloc_device(10 * 77 * 4)
    h0 = b.bla_h2d_bytes()
    b.bla_mnist_gather(store, ptr(idx), 77, xd, yd, 10)
    assert b.bla_h2d_bytes() - h0 == 77 * 4              # only the indices cross PCIe
    X = np.empty((784, 77), np.float32); Y = np.empty((10, 77), np.float32)
    b.bla_copy_d2h(ptr(X), xd, X.nbytes); b.bla_copy_d2h(ptr(Y), yd, Y.nbytes); b.bla_sync()
    assert np.array_equal(X, rows[idx, 1:].T)            # mnist_nn.c:209-211: input_data[k + B*p] = pixel p of sample k
    want_y = np.zeros((10, 77), np.float32); want_y[rows[idx, 0].astype(int), np.arange(77)] = 1
    assert np.array_equal(Y, want_y)
    b.bla_free(xd); b.bla_free(yd); b.bla_mnist_destroy(store)


@pytest.mark.skipif(not P.have("ref_mnist_nn_f64_b512"), reason="oracle/_ref programs not built")
def test_device_epoch_matches_the_reference_program(bla, tmp_path):
    """`mnist_nn init` + `train 3` of the compiled reference (B = 512) against bla_mlp_train_epoch x 3 started from the same
    checkpoint files and the same srand(42) (mnist_nn.c:513): same loss / accuracy curve."""
    b = bla
    d = tmp_path / "ref"
    (d / "data" / "mnist_nn").mkdir(parents=True); (d / "data" / "mnist").mkdir()
    P.mnist_csv(d / "data" / "mnist" / "mnist_train.csv", 1536, 3)
    P.mnist_csv(d / "data" / "mnist" / "mnist_test.csv", 64, 4)
    P.run("ref_mnist_nn_f64_b512", str(d), "init")
    dims = (C.c_int * 4)(784, 256, 128, 10)
    net = b.bla_mlp_create(dims, 512)
    b.bla_mlp_load_csv(net, str(d / "data" / "mnist_nn").encode())          # before `train` rewrites the checkpoint
    want = P.run("ref_mnist_nn_f64_b512", str(d), "train", "3")
    pw = re.findall(r"Epoch (\d+):\s+Avg accuracy: ([0-9.]+)\s+Avg loss: ([0-9.]+)", want)
    assert len(pw) == 3
    store = b.bla_mnist_from_csv(str(d / "data" / "mnist" / "mnist_train.csv").encode())
    libc.srand(42)
    for e, acc, loss in pw:
        stats = np.zeros(2)
        b.bla_mlp_train_epoch(net, store, 512, 0.02, ptr(stats))
        assert abs(stats[0] - float(acc)) <= 2e-3, (e, stats, acc)
        assert abs(stats[1] - float(loss)) <= 1e-4 * max(1.0, float(loss)), (e, stats, loss)
    b.bla_mnist_destroy(store); b.bla_mlp_destroy(net)


def test_cifar_gather_matches_fill_random_data_and_load_example(bla, tmp_path):
    b = bla
    rng = np.random.default_rng(8)
    recs = rng.integers(0, 256, (10000, 3073), dtype=np.uint8)               # lib/cifar10.c:6-11: 10,000 records per batch file
    path = tmp_path / "data_batch_1.bin"
    recs.tofile(path)
    store = b.bla_cifar_open(str(path).encode())
    assert b.bla_cifar_num_examples(store) == 10000
    idx = np.empty(9, np.int32)
    libc.srand(42)
    b.bla_cifar_sample(store, 9, ptr(idx))
    libc.srand(42)
    want_idx = [int(np.float32(np.float32(libc.rand()) / (np.float32(2147483647) + np.float32(1))) * np.float32(10000)) for _ in range(9)]
    assert idx.tolist() == want_idx
    xd = b.bla_malloc_device(9 * 3072 * 4)
    b.bla_cifar_gather(store, ptr(idx), 9, xd)
    got = np.empty((9, 3, 32, 32), np.float32)
    b.bla_copy_d2h(ptr(got), xd, got.nbytes); b.bla_sync()
    pix = recs[idx, 1:].reshape(9, 3, 32, 32)[:, :, ::-1, :].astype(np.float32)     # cifar10.c:24-31: rows bottom-up
    assert np.array_equal(got, ((pix - np.float32(127.5)) / np.float32(127.5)).astype(np.float32))   # cifar_unet.c:229
    ref = ref_data()
    if ref is not None:                                   # the compiled fill_random_data on the same file and rand() stream
        fd = os.open(str(path), os.O_RDONLY)
        buf = np.empty(3072, np.uint8)
        libc.srand(42)
        for k in range(9):
            ref.fill_random_data(fd, buf.ctypes.data_as(C.c_void_p))
            want = ((buf.astype(np.float64) - 127.5) / 127.5).astype(np.float32).reshape(3, 32, 32)
            assert np.array_equal(got[k], want)
        os.close(fd)
    b.bla_free(xd); b.bla_cifar_destroy(store)


def test_sampler_and_gather_edge_cases(bla):
    """a batch of one, a ragged last batch (mnist_nn.c:194-195), more draws than examples (the sampler starts over, mnist_csv2.c:43-46)"""
    b = bla
    n = 37
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, (n, 784)).astype(np.float32); y = rng.integers(0, 10, n).astype(np.float32)
    store = b.bla_mnist_from_arrays(ptr(x), ptr(y), n, 784)
    idx = np.empty(100, np.int32)
    libc.srand(7)
    b.bla_mnist_sample_take(store, 100, ptr(idx))
    assert idx.tolist() == reference_take_sequence(n, 100, 7)
    xd = b.bla_malloc_device(784 * 4); yd = b.bla_malloc_device(10 * 4)
    b.bla_mnist_gather(store, ptr(idx), 1, xd, yd, 10)
    X = np.empty((784, 1), np.float32)
    b.bla_copy_d2h(ptr(X), xd, X.nbytes); b.bla_sync()
    assert np.array_equal(X[:, 0], x[idx[0]])
    b.bla_mnist_gather(store, ptr(idx), 0, xd, yd, 10)          # nothing to do, nothing launched
    dims = (C.c_int * 4)(784, 256, 128, 10)
    net = b.bla_mlp_create(dims, 16)
    b.bla_mlp_init_params(net, 1)
    stats = np.zeros(2)
    b.bla_mlp_train_epoch(net, store, 16, 0.02, ptr(stats))     # 37 = 16 + 16 + 5
    assert 0.0 <= stats[0] <= 1.0 and np.isfinite(stats[1])
    b.bla_free(xd); b.bla_free(yd); b.bla_mlp_destroy(net); b.bla_mnist_destroy(store)


@pytest.mark.parametrize("path", ["fp32", "auto"])
def test_device_hinge_iterations_match_the_oracle(bla, path):
    """BASELINE.json configs[1]: three full-batch iterations of model/mnist_hinge.c:123-166 on the device against the float
    restatement of the same loop (oracle orc_hinge_iter, pinned to the reference program in test_oracle_pinned.py), partial
    gradient clear (:126) included."""
    from helpers import load_oracle, rel_err
    b = bla
    b.bla_set_gemm_path(b.GEMM_FP32 if path == "fp32" else b.GEMM_AUTO)
    o32 = load_oracle(np.float32)
    n, F = 3000, 784
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, (n, F)).astype(np.float32)
    labels = rng.integers(0, 10, n).astype(np.int32)
    w0 = (rng.random((10, F)) / 10 - 0.05).astype(np.float32)           # mnist_hinge.c:20: rand()/(10*RAND_MAX) - 0.05
    store = b.bla_mnist_from_arrays(ptr(x), ptr(labels.astype(np.float32)), n, F)
    h = b.bla_hinge_create(F, 10, n)
    b.bla_hinge_set_weights(h, ptr(w0))
    w_ref, grad_ref = w0.copy(), np.zeros((10, F), np.float32)
    try:
        for it in range(3):
            norms = np.zeros(10, np.float32); norms_ref = np.zeros(10, np.float32)
            b.bla_hinge_iteration(h, store, 0.001, ptr(norms))
            o32.orc_hinge_iter(n, F, ptr(w_ref), ptr(grad_ref), ptr(x), ptr(labels), C.c_float(0.001), ptr(norms_ref))
            assert np.allclose(norms, norms_ref, rtol=1e-4, atol=1e-6), (it, norms, norms_ref)
        got = np.empty_like(w0)
        b.bla_hinge_get_weights(h, ptr(got))
        assert rel_err(got, w_ref) <= 1e-4, rel_err(got, w_ref)
    finally:
        b.bla_hinge_destroy(h); b.bla_mnist_destroy(store)
        b.bla_set_gemm_path(b.GEMM_FP32)
