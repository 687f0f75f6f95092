"""One GEMM shape, a few launches -- the target of `ncu --set full` captures.
Usage: python profiles/prof_one.py ta tb M N K [path] [int_b]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b

ta, tb, M, N, K = [int(v) for v in sys.argv[1:6]]
path = {"fp32": b.GEMM_FP32, "3xtf32": b.GEMM_3XTF32}[sys.argv[6] if len(sys.argv) > 6 else "3xtf32"]
int_b = len(sys.argv) > 7 and sys.argv[7] == "1"
b.bla_init(0)
b.bla_set_gemm_path(path)
A = b.bla_malloc_device(M * K * 4); B = b.bla_malloc_device(K * N * 4); Cm = b.bla_malloc_device(M * N * 4)
b.bla_fill_uniform(A, M * K, 1, -0.5, 0.5)
if int_b:
    px = np.random.default_rng(0).integers(0, 256, K * N, dtype=np.uint8)
    pd = b.bla_malloc_device(px.nbytes)
    b.bla_copy_h2d(pd, px.ctypes.data_as(C.c_void_p), px.nbytes)
    b.bla_u8_to_float(B, pd, px.size, 1.0)
else:
    b.bla_fill_uniform(B, K * N, 2, -0.5, 0.5)
for _ in range(5):
    b.bla_gemm(ta, tb, M, N, K, A, M if ta else K, B, K if tb else N, Cm, N)
b.bla_sync()
print("done")
