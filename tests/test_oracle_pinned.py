"""Pins the CPU restatement (oracle/bla_oracle.c) to the reference: (1) against the committed golden
vectors generated from the reference's own compiled C (tests/golden/make_golden.py), (2) directly
against oracle/_ref/libref_*.so when it is present (it travels to the GPU box).  Bit-exact: the
restatement keeps the reference's accumulation order."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, as_matrix, load_oracle, load_ref, matrix_to_numpy, ptr, ref_available

DTYPES = {"f64": np.float64, "f32": np.float32}


def golden(variant):
    return np.load(os.path.join(GOLDEN_DIR, f"ref_{variant}.npz"))


@pytest.mark.parametrize("variant", ["f64", "f32"])
def test_matrix_ops_match_golden(variant):
    g, dt = golden(variant), DTYPES[variant]
    o = load_oracle(dt)
    a, b = g["gemm_a"], g["gemm_b"]
    c = np.empty((a.shape[0], b.shape[1]), dt)
    o.orc_gemm(a.shape[0], a.shape[1], b.shape[1], ptr(a), ptr(b), ptr(c))
    assert np.array_equal(c, g["gemm_c"])
    ka, kb = g["kat_a"], g["kat_b"]
    kc = np.empty((2, 2), dt)
    o.orc_gemm(2, 3, 2, ptr(ka), ptr(kb), ptr(kc))
    assert np.array_equal(kc, g["kat_c"])
    np.testing.assert_allclose(kc, [[1.4, 8.5], [5.0, 19.0]], rtol=1e-6)      # main.c:20-41

    x, y = g["ew_x"], g["ew_y"]
    R, Cc = x.shape
    t = x.copy(); o.orc_scale(t.size, ptr(t), 1 / np.float32(255.0)); assert np.array_equal(t, g["scale"])
    t = x.copy(); o.orc_add(C.c_size_t(t.size), ptr(t), ptr(y)); assert np.array_equal(t, g["add"])
    t = x.copy(); o.orc_hadamard(C.c_size_t(t.size), ptr(t), ptr(y)); assert np.array_equal(t, g["hadamard"])
    t = x.copy(); o.orc_transpose(R, Cc, ptr(t)); assert np.array_equal(t.reshape(Cc, R), g["transpose"])
    assert np.array_equal(g["transpose"], x.T)
    out = np.empty((1, Cc), dt); o.orc_row_sum(R, Cc, ptr(x), ptr(out)); assert np.array_equal(out, g["row_sum"])
    out = np.empty((R, 1), dt); o.orc_col_sum(R, Cc, ptr(x), ptr(out), 1); assert np.array_equal(out, g["col_sum"])
    # the D2 quirk really is a window sum over the flat buffer, not a row total
    flat = x.ravel()
    want = np.array([flat[i * R:i * R + Cc].astype(np.float64).sum() for i in range(R)])
    np.testing.assert_allclose(out.ravel(), want, rtol=1e-5)
    assert o.orc_frobenius(R, Cc, ptr(x)) == g["frobenius"]
    assert o.orc_max(C.c_size_t(x.size), ptr(x)) == g["max_value"]
    t = x.copy(); o.orc_zscore(C.c_size_t(t.size), ptr(t)); assert np.array_equal(t, g["zscore"])
    t = x.copy(); o.orc_add_tile_columns(R, Cc, ptr(t), 1, ptr(g["tile_cols_b"])); assert np.array_equal(t, g["tile_cols"])
    t = x.copy(); o.orc_add_tile_columns(R, Cc, ptr(t), 3, ptr(g["tile_cols3_b"])); assert np.array_equal(t, g["tile_cols3"])
    t = x.copy(); o.orc_add_tile_rows(R, Cc, ptr(t), ptr(g["tile_rows_b"])); assert np.array_equal(t, g["tile_rows"])
    t = x.copy(); o.orc_relu(C.c_size_t(t.size), ptr(t)); assert np.array_equal(t, g["relu"])
    t = (4 * x).copy(); o.orc_softmax_cols(R, Cc, ptr(t)); assert np.array_equal(t, g["softmax_cols"])
    t = (4 * x).copy(); o.orc_softmax_rows(R, Cc, ptr(t)); assert np.array_equal(t, g["softmax_rows"])


@pytest.mark.parametrize("variant", ["f64", "f32"])
@pytest.mark.parametrize("case", ["s1k3", "s2k3", "s1k1", "s1k3_ragged"])
def test_conv_matches_golden(variant, case):
    g, dt = golden(variant), DTYPES[variant]
    o = load_oracle(dt)
    Cin, H, W, F, k, s = [int(v) for v in g[f"conv_{case}_meta"]]
    x, kr = g[f"conv_{case}_x"], g[f"conv_{case}_k"]
    Ho, Wo = -(-H // s), -(-W // s)
    col = np.empty((Ho * Wo, Cin * k * k), dt)
    o.orc_im2col(Cin, H, W, k, s, ptr(x), ptr(col))
    assert np.array_equal(col, g[f"conv_{case}_im2col"])
    y = np.empty((F, Ho, Wo), dt)
    o.orc_conv(Cin, H, W, F, k, s, ptr(x), ptr(kr), ptr(y))
    assert np.array_equal(y, g[f"conv_{case}_y"])
    if f"conv_{case}_dy" in g.files:
        dk, dx = np.empty_like(kr), np.empty_like(x)
        o.orc_conv_ddx(Cin, H, W, F, k, s, ptr(x), ptr(kr), ptr(g[f"conv_{case}_dy"]), ptr(dk), ptr(dx))
        assert np.array_equal(dk, g[f"conv_{case}_dk"])
        assert np.array_equal(dx, g[f"conv_{case}_dx"])


def test_conv_stride2_dgrad_is_the_adjoint():
    """the reference has no valid stride-2 dgrad, so the restatement's is validated as the
    exact adjoint of its own (golden-pinned) forward: <conv(x), dy> == <x, dx> and == <k, dk>."""
    o = load_oracle(np.float64)
    rng = np.random.default_rng(3)
    Cin, H, W, F, k, s = 3, 9, 8, 4, 3, 2
    Ho, Wo = -(-H // s), -(-W // s)
    x = rng.normal(size=(Cin, H, W)); kr = rng.normal(size=(F, Cin, k, k)); dy = rng.normal(size=(F, Ho, Wo))
    y = np.empty((F, Ho, Wo)); dk = np.empty_like(kr); dx = np.empty_like(x)
    o.orc_conv(Cin, H, W, F, k, s, ptr(x), ptr(kr), ptr(y))
    o.orc_conv_ddx(Cin, H, W, F, k, s, ptr(x), ptr(kr), ptr(dy), ptr(dk), ptr(dx))
    np.testing.assert_allclose((y * dy).sum(), (x * dx).sum(), rtol=1e-12)
    np.testing.assert_allclose((y * dy).sum(), (kr * dk).sum(), rtol=1e-12)


@pytest.mark.parametrize("variant", ["f64", "f32"])
@pytest.mark.parametrize("case", ["even", "ragged", "rgb"])
def test_group_norm_matches_golden(variant, case):
    g, dt = golden(variant), DTYPES[variant]
    o = load_oracle(dt)
    Cn, H, W, gs = [int(v) for v in g[f"gn_{case}_meta"]]
    G = -(-Cn // gs)
    x = g[f"gn_{case}_x"]
    y = np.empty_like(x); var = np.empty(G, dt); mu = np.empty(G, dt)
    o.orc_group_norm(Cn, H * W, gs, ptr(x), ptr(y), ptr(var), ptr(mu), 1)
    assert np.array_equal(y, g[f"gn_{case}_y"])
    assert np.array_equal(var, g[f"gn_{case}_var"]) and np.array_equal(mu, g[f"gn_{case}_mean"])
    # D5: the stored "stdev" is the variance and the output is divided by it
    np.testing.assert_allclose(var[0], x[:gs].astype(np.float64).var(), rtol=1e-5)
    dx = np.empty_like(x)
    o.orc_group_norm_ddx(Cn, H * W, gs, ptr(g[f"gn_{case}_dy"]), ptr(dx), ptr(x), ptr(mu), ptr(var))
    assert np.array_equal(dx, g[f"gn_{case}_dx"])


@pytest.mark.parametrize("tag,sizes,act,param", [("main", [3, 2, 2], 2, 0.1), ("mfm", [2, 3, 2], 1, 0.0)])
def test_layer_matches_golden(tag, sizes, act, param):
    """lib/layer.c through the float build; `main` is the reference's own smoke program (main.c:52-83):
    nodes ~ [2.47, 5.39], W' ~ [[0.91,1.80],[2.78,3.51]], b' ~ [0.08,0.15]."""
    g = np.load(os.path.join(GOLDEN_DIR, "ref_layer_f32.npz"))
    o = load_oracle(np.float32)
    L = len(sizes) - 1
    Ws = [g[f"{tag}_W{i}_in"].copy() for i in range(1, L + 1)]
    bs = [g[f"{tag}_b{i}_in"].copy().reshape(-1) for i in range(1, L + 1)]
    x = g[f"{tag}_x"].copy()
    raws = [np.empty(sizes[i], np.float32) for i in range(1, L + 1)]
    nodes = [np.empty(sizes[i], np.float32) for i in range(1, L + 1)]
    prev = x
    for i in range(L):
        o.orc_dense_forward(sizes[i + 1], sizes[i], ptr(Ws[i]), ptr(bs[i]), ptr(prev), act, param, ptr(raws[i]), ptr(nodes[i]))
        assert np.array_equal(raws[i], g[f"{tag}_raw{i + 1}"].ravel())
        assert np.array_equal(nodes[i], g[f"{tag}_nodes{i + 1}"].ravel())
        prev = nodes[i]
    arr = lambda lst: (C.c_void_p * L)(*[a.ctypes.data for a in lst])
    szs = (C.c_int * (L + 1))(*sizes)
    acts = (C.c_int * L)(*([act] * L))
    params = (C.c_double * L)(*([param] * L))
    o.orc_dense_backprop.argtypes = [C.c_int, C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
    o.orc_dense_backprop(L, szs, arr(Ws), arr(bs), arr(raws), arr(nodes), ptr(x), acts, params, ptr(g[f"{tag}_target"]),
                         float(g[f"{tag}_lr"]))
    for i in range(L):
        assert np.array_equal(Ws[i], g[f"{tag}_W{i + 1}_out"])
        assert np.array_equal(bs[i], g[f"{tag}_b{i + 1}_out"].ravel())
    if tag == "main":
        np.testing.assert_allclose(nodes[1], [2.475, 5.391], atol=1e-3)
        np.testing.assert_allclose(Ws[1].ravel(), [0.9129, 1.8001, 2.7843, 3.5050], atol=1e-4)


@pytest.mark.skipif(not ref_available("f64"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("variant", ["f64", "f32"])
def test_restatement_matches_live_reference(variant):
    """Fresh random shapes straight against the compiled reference, beyond the committed fixtures."""
    dt = DTYPES[variant]
    ref, o = load_ref(variant), load_oracle(dt)
    rng = np.random.default_rng(99)
    for (M, K, N) in [(1, 1, 1), (5, 1, 7), (64, 96, 33), (10, 128, 200), (1, 784, 1)]:
        a = rng.normal(size=(M, K)).astype(dt); b = rng.normal(size=(K, N)).astype(dt)
        want = matrix_to_numpy(ref.matrix_multiply(as_matrix(ref, a), as_matrix(ref, b)), dt)
        got = np.empty((M, N), dt)
        o.orc_gemm(M, K, N, ptr(a), ptr(b), ptr(got))
        assert np.array_equal(got, want)
    for (R, Cc) in [(1, 1), (3, 3), (10, 64), (128, 512)]:
        x = rng.normal(size=(R, Cc)).astype(dt)
        want = matrix_to_numpy(ref.matrix_col_sum(as_matrix(ref, x)), dt)
        got = np.empty((R, 1), dt); o.orc_col_sum(R, Cc, ptr(x), ptr(got), 1)
        assert np.array_equal(got, want)
        assert o.orc_frobenius(R, Cc, ptr(x)) == ref.frobenius_norm(as_matrix(ref, x))
