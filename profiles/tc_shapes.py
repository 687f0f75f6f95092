"""Micro-benchmark of the GEMM paths on the MLP's shapes and layouts (device-resident, CUDA events on the
library stream).  Usage: python profiles/tc_shapes.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bla_b200 as b

b.bla_init(0)
stream = torch.cuda.Stream()
b.bla_set_stream(C.c_void_p(stream.cuda_stream))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


CASES = [("fwd1   NN 256x60000x784", 0, 0, 256, 60000, 784), ("fwd1   NT (B stored [n][k])", 0, 1, 256, 60000, 784),
         ("fwd1-ish NN 256x60032x768", 0, 0, 256, 60032, 768),
         ("longK  NN 256x60000x4096", 0, 0, 256, 60000, 4096), ("square-ish NN 8192x8192x784", 0, 0, 8192, 8192, 784),
         ("wgrad1 NT 256x784x60000", 0, 1, 256, 784, 60000), ("wgrad1 NT 256x768x61440", 0, 1, 256, 768, 61440),
         ("dgrad2 TN 256x60000x128", 1, 0, 256, 60000, 128),
         ("fwd2   NN 128x60000x256", 0, 0, 128, 60000, 256), ("wgrad2 NT 128x256x60000", 0, 1, 128, 256, 60000),
         ("square NN 4096", 0, 0, 4096, 4096, 4096), ("square NT 4096", 0, 1, 4096, 4096, 4096), ("square TN 4096", 1, 0, 4096, 4096, 4096)]
for name, ta, tb, M, N, K in CASES:
    A = b.bla_malloc_device(M * K * 4); B = b.bla_malloc_device(K * N * 4); Cm = b.bla_malloc_device(M * N * 4)
    b.bla_fill_uniform(A, M * K, 1, -0.5, 0.5); b.bla_fill_uniform(B, K * N, 2, -0.5, 0.5)
    lda = M if ta else K
    ldb = K if tb else N
    out = []
    for path in (b.GEMM_3XTF32, b.GEMM_FP32):
        b.bla_set_gemm_path(path)
        ms = timeit(lambda: b.bla_gemm(ta, tb, M, N, K, A, lda, B, ldb, Cm, N))
        out.append(f"{2.0 * M * N * K / ms / 1e9:7.1f} TF/s ({ms * 1e3:7.1f} us)")
    print(f"{name:34s} 3xTF32 {out[0]}   FP32 {out[1]}", flush=True)
    for p_ in (A, B, Cm):
        b.bla_free(p_)
