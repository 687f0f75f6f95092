"""-m gpu: the HEADLINE sizes against the compiled reference.  tests/golden/headline.npz holds what the
reference's own C (oracle/_ref/libref_f64.so, driven by tests/golden/make_golden_headline.py) produced for
  * one SGD step of the MNIST MLP on a 60,000-column batch (BASELINE.json configs[2]; model/mnist_nn.c:218-315), and
  * matrix_multiply at 1024^2 and 2048^2 (configs[3]; lib/matrix.c:35-57; SURVEY 8(d) config 4: "parity vs ref_f64 for N <= 2048"),
on seeded inputs that tests/headline_inputs.py regenerates here.  Tolerances are BASELINE.json's: FP32 path <= 1e-5, 3xTF32 <= 1e-3
(norm-wise relative), same loss within tolerance, same number of argmax hits up to near-ties."""
import ctypes as C
import os

import numpy as np
import pytest

from headline_inputs import DIMS, gemm_inputs, mlp_batch, mlp_params, sample_index, splitmix_uniform
from helpers import GOLDEN_DIR, ptr, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(GOLDEN_DIR, "headline.npz")


@pytest.fixture(scope="module")
def bla():
    import bla_b200 as b
    if b.bla_device_count() < 1:
        pytest.skip("no CUDA device")
    b.bla_init(0)
    yield b
    b.bla_set_gemm_path(b.GEMM_AUTO)


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/headline.npz missing")
@pytest.mark.parametrize("path,tol", [("fp32", 1e-5), ("3xtf32", 1e-3)])
def test_mlp_step_at_60000_columns_vs_the_compiled_reference(bla, path, tol):
    b = bla
    gold = np.load(GOLD)
    b.bla_set_gemm_path(b.GEMM_FP32 if path == "fp32" else b.GEMM_3XTF32)
    B = 60000
    X, Y = mlp_batch(B)
    p0 = mlp_params()
    dims = (C.c_int * 4)(*DIMS)
    net = b.bla_mlp_create(dims, B)
    try:
        b.bla_mlp_set_params(net, *[ptr(np.ascontiguousarray(p)) for p in p0])
        stats = np.zeros(2)
        tc0 = b.bla_tc_launch_count()
        b.bla_mlp_set_host_chunking(net, 0)
        b.bla_mlp_train_step(net, ptr(X), ptr(Y), B, B, 0, 0.02, ptr(stats))
        if path == "3xtf32":
            assert b.bla_tc_launch_count() > tc0          # the tensor path really ran
        got = [np.empty_like(p) for p in p0]
        b.bla_mlp_get_params(net, *[ptr(g) for g in got])
    finally:
        b.bla_mlp_destroy(net)
    assert abs(stats[0] - float(gold["mlp_loss"])) <= max(tol, 1e-6) * abs(float(gold["mlp_loss"])), (stats[0], float(gold["mlp_loss"]))
    assert abs(int(stats[1]) - int(gold["mlp_correct"])) <= 3, (stats[1], int(gold["mlp_correct"]))    # near-ties of the argmax
    for i, (g, p) in enumerate(zip(got, p0)):
        flat = g.ravel().astype(np.float64)
        idx = sample_index(flat.size, 4096, 100 + i)
        want = gold[f"mlp_p{i}_sample"]
        upd_got, upd_want = flat[idx] - p.ravel()[idx].astype(np.float64), want - p.ravel()[idx].astype(np.float64)
        assert rel_err(flat[idx], want) <= tol, (i, rel_err(flat[idx], want))
        assert rel_err(upd_got, upd_want) <= 5 * tol, (i, "update", rel_err(upd_got, upd_want))       # the update alone (no help from p0)
        assert abs(np.linalg.norm(flat) - float(gold[f"mlp_p{i}_norm"])) <= tol * float(gold[f"mlp_p{i}_norm"]), i


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/headline.npz missing")
@pytest.mark.parametrize("n", [1024, 2048])
def test_square_gemm_vs_the_compiled_reference(bla, n):
    b = bla
    gold = np.load(GOLD)
    A, Bm = gemm_inputs(n)
    Ad, Bd, Cd = (b.bla_malloc_device(n * n * 4) for _ in range(3))
    try:
        # the device generator the bench sweep uses produces exactly these operands
        b.bla_fill_uniform(Ad, n * n, 1, -0.5, 0.5); b.bla_fill_uniform(Bd, n * n, 2, -0.5, 0.5)
        chk = np.empty((n, n), np.float32)
        b.bla_copy_d2h(ptr(chk), Ad, chk.nbytes); b.bla_sync()
        assert np.array_equal(chk, A)
        b.bla_copy_d2h(ptr(chk), Bd, chk.nbytes); b.bla_sync()
        assert np.array_equal(chk, Bm)
        v = np.random.default_rng(7).uniform(-1, 1, n)
        idx = sample_index(n * n, 4096, 200 + n)
        for path, tol in ((b.GEMM_FP32, 1e-5), (b.GEMM_3XTF32, 1e-3)):
            b.bla_set_gemm_path(path)
            tc0 = b.bla_tc_launch_count()
            b.bla_gemm(0, 0, n, n, n, Ad, n, Bd, n, Cd, n)
            out = np.empty((n, n), np.float32)
            b.bla_copy_d2h(ptr(out), Cd, out.nbytes); b.bla_sync()
            if path == b.GEMM_3XTF32:
                assert b.bla_tc_launch_count() > tc0
            o64 = out.astype(np.float64)
            assert rel_err(o64.ravel()[idx], gold[f"gemm{n}_sample"]) <= tol, (path, rel_err(o64.ravel()[idx], gold[f"gemm{n}_sample"]))
            assert rel_err(o64 @ v, gold[f"gemm{n}_Cv"]) <= tol and rel_err(v @ o64, gold[f"gemm{n}_vC"]) <= tol
            assert abs(np.linalg.norm(o64) - float(gold[f"gemm{n}_norm"])) <= tol * float(gold[f"gemm{n}_norm"])
    finally:
        for p_ in (Ad, Bd, Cd):
            b.bla_free(p_)
        b.bla_set_gemm_path(b.GEMM_AUTO)


def test_splitmix_twin_is_deterministic():
    a = splitmix_uniform(8, 1, -0.5, 0.5)
    assert a.dtype == np.float32 and np.all(np.abs(a) <= 0.5) and not np.array_equal(a, splitmix_uniform(8, 2, -0.5, 0.5))
