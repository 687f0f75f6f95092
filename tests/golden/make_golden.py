#!/usr/bin/env python
"""Generate the committed golden vectors from the REAL reference (oracle/_ref/libref_*.so, built by
oracle/build_ref.sh from /root/reference).  Run in the build container only:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Every array in tests/golden/*.npz is an input or an output of the reference's own compiled C on
that input; nothing here comes from the restatement or from the CUDA product.  The reference
itself ships no golden files; its only known-answer material is main.c, which is
reproduced in `kat_main`."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import (GOLDEN_DIR, as_matrix, kernel_table, load_ref, matrix_to_numpy, planes)  # noqa: E402


def gen_matrix(ref, dt, out):
    rng = np.random.default_rng(1234)
    P = C.POINTER(ref.MatrixT)
    a = rng.uniform(-0.5, 0.5, (37, 53)).astype(dt)
    b = rng.uniform(-0.5, 0.5, (53, 29)).astype(dt)
    c = ref.matrix_multiply(as_matrix(ref, a), as_matrix(ref, b))
    out.update(gemm_a=a, gemm_b=b, gemm_c=matrix_to_numpy(c, dt))
    # main.c:20-41 known-answer GEMM
    ka = np.array([[1, 2, 3], [4, 5, 6]], dtype=dt)
    kb = np.array([[1, 0.5], [0.2, 1], [0, 2]], dtype=dt)
    out.update(kat_a=ka, kat_b=kb, kat_c=matrix_to_numpy(ref.matrix_multiply(as_matrix(ref, ka), as_matrix(ref, kb)), dt))

    x = rng.normal(0, 1, (13, 40)).astype(dt)   # cols >= rows: the D2 quirk is well defined
    y = rng.normal(0, 1, (13, 40)).astype(dt)
    out.update(ew_x=x, ew_y=y)
    t = x.copy(); m = as_matrix(ref, t); ref.matrix_scale(C.byref(m), ref.real(1 / np.float32(255.0))); out["scale"] = t
    t = x.copy(); m = as_matrix(ref, t); n = as_matrix(ref, y); ref.matrix_add(C.byref(m), C.byref(n)); out["add"] = t
    t = x.copy(); m = as_matrix(ref, t); ref.matrix_multiply_elementwise(C.byref(m), C.byref(n)); out["hadamard"] = t
    t = x.copy(); m = as_matrix(ref, t); ref.matrix_transpose(C.byref(m)); out["transpose"] = t.reshape(40, 13).copy()
    assert (m.rows, m.cols) == (40, 13)
    out["row_sum"] = matrix_to_numpy(ref.matrix_row_sum(as_matrix(ref, x)), dt)
    out["col_sum"] = matrix_to_numpy(ref.matrix_col_sum(as_matrix(ref, x)), dt)
    out["frobenius"] = np.array(ref.frobenius_norm(as_matrix(ref, x)), dtype=dt)
    out["max_value"] = np.array(ref.max_value(as_matrix(ref, x)), dtype=dt)
    t = x.copy(); m = as_matrix(ref, t); ref.matrix_z_score_normalize(C.byref(m)); out["zscore"] = t
    bias = rng.normal(0, 1, (13, 1)).astype(dt)
    t = x.copy(); m = as_matrix(ref, t); bm = as_matrix(ref, bias); ref.matrix_add_tile_columns(C.byref(m), C.byref(bm))
    out.update(tile_cols_b=bias, tile_cols=t)
    bias3 = rng.normal(0, 1, (13, 3)).astype(dt)
    t = x.copy(); m = as_matrix(ref, t); bm = as_matrix(ref, bias3); ref.matrix_add_tile_columns(C.byref(m), C.byref(bm))
    out.update(tile_cols3_b=bias3, tile_cols3=t)
    rb = rng.normal(0, 1, (1, 40)).astype(dt)
    t = x.copy(); m = as_matrix(ref, t); bm = as_matrix(ref, rb); ref.matrix_add_tile_rows(C.byref(m), C.byref(bm))
    out.update(tile_rows_b=rb, tile_rows=t)
    # lib/util.c activations
    t = x.copy(); ref.relu(t.ctypes.data_as(C.c_void_p), C.c_int(t.size)); out["relu"] = t
    t = (4 * x).copy(); ref.softmax(t.ctypes.data_as(C.c_void_p), 13, 40); out["softmax_cols"] = t
    t = (4 * x).copy(); ref.softmax_row_wise(t.ctypes.data_as(C.c_void_p), 13, 40); out["softmax_rows"] = t
    _ = P


def run_conv(ref, dt, Cin, H, W, F, k, s, seed, with_ddx):
    rng = np.random.default_rng(seed)
    M = ref.MatrixT
    Ho, Wo = -(-H // s), -(-W // s)
    x = rng.normal(0, 1, (Cin, H, W)).astype(dt)
    kr = rng.normal(0, 0.3, (F, Cin, k, k)).astype(dt)

    def conv_data():
        bufs = dict(im2col=np.zeros((Ho * Wo, k * k * Cin), dt), kernel_matrix=np.zeros((k * k * Cin, F), dt),
                    product=np.zeros((Ho * Wo, F), dt), output=np.zeros((F, Ho, Wo), dt))
        mats = {n: M(v.shape[0], v.shape[1], v.ctypes.data_as(C.POINTER(ref.real))) for n, v in bufs.items() if n != "output"}
        outp = planes(ref, bufs["output"])
        cd = ref.ConvDataT(C.pointer(mats["im2col"]), C.pointer(mats["kernel_matrix"]), C.pointer(mats["product"]),
                           C.cast(outp, C.POINTER(M)))
        cd._keep = (bufs, mats, outp)
        return cd, bufs

    cd, bufs = conv_data()
    xp = planes(ref, x)
    kt = kernel_table(ref, kr)
    ref.conv(C.cast(xp, C.POINTER(M)), kt, C.byref(cd), Cin, F, s)
    res = dict(x=x, k=kr, y=bufs["output"].copy(), im2col=bufs["im2col"].copy(), meta=np.array([Cin, H, W, F, k, s]))
    if with_ddx:
        dy = rng.normal(0, 1, (F, Ho, Wo)).astype(dt)
        gcd, gbufs = conv_data()
        dk = np.zeros_like(kr)
        dx = np.zeros_like(x)
        dyp = planes(ref, dy); dkt = kernel_table(ref, dk); dxp = planes(ref, dx)
        ref.conv_ddx(C.cast(dyp, C.POINTER(M)), C.byref(cd), C.byref(gcd), dkt, C.cast(dxp, C.POINTER(M)), Cin, s)
        res.update(dy=dy, dk=dk, dx=dx)
    return res


def gen_conv(ref, dt, out):
    cases = {"s1k3": (3, 8, 8, 4, 3, 1, True), "s2k3": (4, 9, 7, 5, 3, 2, False), "s1k1": (5, 6, 6, 3, 1, 1, True),
             "s1k3_ragged": (2, 5, 7, 3, 3, 1, True)}
    for i, (name, (Cin, H, W, F, k, s, ddx)) in enumerate(cases.items()):
        for key, v in run_conv(ref, dt, Cin, H, W, F, k, s, 100 + i, ddx).items():
            out[f"conv_{name}_{key}"] = v


def gen_norm(ref, dt, out):
    rng = np.random.default_rng(77)
    M = ref.MatrixT
    for name, (Cn, H, W, gs) in {"even": (8, 4, 4, 4), "ragged": (6, 5, 5, 4), "rgb": (3, 8, 8, 32)}.items():
        G = -(-Cn // gs)
        x = rng.normal(0.3, 1.5, (Cn, H, W)).astype(dt)
        y = np.zeros_like(x); sd = np.zeros(G, dt); mu = np.zeros(G, dt)
        ref.group_norm(C.cast(planes(ref, x), C.POINTER(M)), C.cast(planes(ref, y), C.POINTER(M)),
                       sd.ctypes.data_as(C.c_void_p), mu.ctypes.data_as(C.c_void_p), Cn, gs)
        dy = rng.normal(0, 1, (Cn, H, W)).astype(dt)
        dx = np.zeros_like(x)
        ref.group_norm_ddx(C.cast(planes(ref, dy), C.POINTER(M)), C.cast(planes(ref, dx), C.POINTER(M)),
                           C.cast(planes(ref, x), C.POINTER(M)), mu.ctypes.data_as(C.c_void_p),
                           sd.ctypes.data_as(C.c_void_p), Cn, gs)
        out.update({f"gn_{name}_x": x, f"gn_{name}_y": y, f"gn_{name}_var": sd, f"gn_{name}_mean": mu,
                    f"gn_{name}_dy": dy, f"gn_{name}_dx": dx, f"gn_{name}_meta": np.array([Cn, H, W, gs])})


class LayerF32(C.Structure):
    pass


ACT = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int)


def gen_layer(out):
    """main.c:52-83 (3-2-2 net, activation 0.1x, target .5/.5, lr .05) and the my_first_model 2-3-2
    ReLU net, through the float build of lib/layer.c."""
    ref = load_ref("f32")
    M = ref.MatrixT
    P = C.POINTER(M)
    LayerF32._fields_ = [("num_nodes", C.c_int), ("nodes", P), ("raw_nodes", P), ("weights", P), ("biases", P),
                         ("previous_layer", C.POINTER(LayerF32)), ("activation", ACT), ("activation_ddx", ACT),
                         ("has_previous_layer", C.c_char), ("has_nodes", C.c_char)]
    ref.back_propagate_errors.argtypes = [C.POINTER(LayerF32), C.POINTER(C.c_float), C.c_float]

    def scale_act(d, n):
        for i in range(n):
            d[i] = np.float32(np.float64(d[i]) * 0.1)   # data[i] *= 0.1 (double constant), main.c:9

    def scale_ddx(d, n):
        for i in range(n):
            d[i] = 0.1

    def relu(d, n):
        for i in range(n):
            if d[i] < 0:
                d[i] = 0

    def relu_ddx(d, n):
        for i in range(n):
            d[i] = 1.0 if d[i] > 0 else 0.0

    def net(sizes, Ws, bs, x, act, ddx, target, lr, tag):
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p

        def heap(arr):  # layer.c frees with libc free(): give it malloc'd memory
            arr = np.ascontiguousarray(arr, np.float32)
            p = libc.malloc(C.c_size_t(arr.nbytes))
            C.memmove(p, arr.ctypes.data, arr.nbytes)
            return C.cast(p, C.POINTER(C.c_float))

        layers = [LayerF32() for _ in sizes]
        layers[0].num_nodes = sizes[0]
        layers[0].nodes = ref.make_matrix(sizes[0], 1, heap(x))
        layers[0].has_nodes = b"\x01"; layers[0].has_previous_layer = b"\x00"
        fa, fd = ACT(act), ACT(ddx)
        for i in range(1, len(sizes)):
            L = layers[i]
            L.num_nodes = sizes[i]
            L.weights = ref.make_matrix(sizes[i], sizes[i - 1], heap(Ws[i - 1]))
            L.biases = ref.make_matrix(sizes[i], 1, heap(bs[i - 1]))
            L.previous_layer = C.pointer(layers[i - 1])
            L.activation, L.activation_ddx = fa, fd
            L.has_previous_layer = b"\x01"; L.has_nodes = b"\x00"
        for i in range(1, len(sizes)):
            ref.feed_forward(C.byref(layers[i]))
        for i in range(1, len(sizes)):
            out[f"{tag}_raw{i}"] = matrix_to_numpy(layers[i].raw_nodes, np.float32)
            out[f"{tag}_nodes{i}"] = matrix_to_numpy(layers[i].nodes, np.float32)
        t = np.asarray(target, np.float32)
        ref.back_propagate_errors(C.byref(layers[-1]), t.ctypes.data_as(C.POINTER(C.c_float)), C.c_float(lr))
        for i in range(1, len(sizes)):
            out[f"{tag}_W{i}_in"] = np.asarray(Ws[i - 1], np.float32)
            out[f"{tag}_b{i}_in"] = np.asarray(bs[i - 1], np.float32)
            out[f"{tag}_W{i}_out"] = matrix_to_numpy(layers[i].weights, np.float32)
            out[f"{tag}_b{i}_out"] = matrix_to_numpy(layers[i].biases, np.float32)
        out[f"{tag}_x"] = np.asarray(x, np.float32)
        out[f"{tag}_target"] = t
        out[f"{tag}_lr"] = np.float32(lr)

    # main.c: data/inputs.csv = 3,7,9 ; weights.csv = 1..6 ; biases.csv = .1,.2 ; both layers load the
    # same files (the 2x2 output layer takes the first four weights)
    net([3, 2, 2], [np.array([[1, 2, 3], [4, 5, 6]]), np.array([[1, 2], [3, 4]])], [np.array([.1, .2]), np.array([.1, .2])],
        np.array([3, 7, 9]), scale_act, scale_ddx, [0.5, 0.5], 0.05, "main")
    rng = np.random.default_rng(5)
    net([2, 3, 2], [rng.normal(0, 1, (3, 2)), rng.normal(0, 1, (2, 3))], [rng.normal(0, .5, 3), rng.normal(0, .5, 2)],
        np.array([0.7, -0.3]), relu, relu_ddx, [1, 0], 0.01, "mfm")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for variant, dt in (("f64", np.float64), ("f32", np.float32)):
        out = {}
        gen_matrix(load_ref(variant), dt, out)
        gen_norm(load_ref(variant), dt, out)
        gen_conv(load_ref(variant + "_convfix"), dt, out)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"ref_{variant}.npz"), **out)
        print(variant, len(out), "arrays")
    out = {}
    gen_layer(out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_layer_f32.npz"), **out)
    print("layer", len(out), "arrays; main.c KAT nodes:", out["main_nodes2"].ravel(), "W':", out["main_W2_out"].ravel(),
          "b':", out["main_b2_out"].ravel())


if __name__ == "__main__":
    main()
