"""Accuracy of the tensor-path GEMM on sub-wave shapes (narrow tiles, profiles of the MLP's data-parallel shards) against float64.
Usage: python profiles/narrow_check.py   (BLA_TC_NARROW=0 for the wide-tile / split-K choice)"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bla_b200 as b
from helpers import ptr, rel_err
b.bla_init(0)
b.bla_set_gemm_path(b.GEMM_3XTF32)
rng = np.random.default_rng(0)
def dev(a):
    d = b.bla_malloc_device(a.nbytes); b.bla_copy_h2d(d, ptr(a), a.nbytes); b.bla_sync(); return d
for (ta, tb, M, N, K) in ((0, 0, 256, 256, 784), (0, 0, 256, 3000, 784), (0, 0, 256, 7500, 784), (0, 0, 128, 3000, 256), (0, 0, 128, 7500, 256),
                          (1, 0, 256, 3000, 128), (1, 0, 256, 7500, 128), (0, 1, 256, 784, 256), (0, 1, 256, 784, 3000), (0, 1, 128, 256, 3000),
                          (0, 0, 128, 1024, 256), (1, 0, 256, 256, 128), (0, 0, 128, 256, 256)):
    A = rng.uniform(-0.08, 0.08, (M, K)).astype(np.float32)
    Bm = rng.uniform(-0.5, 0.5, (K, N)).astype(np.float32)
    want = A.astype(np.float64) @ Bm.astype(np.float64)
    As = np.ascontiguousarray(A.T) if ta else A
    Bs = np.ascontiguousarray(Bm.T) if tb else Bm
    Ad, Bd = dev(As), dev(Bs); Cd = b.bla_malloc_device(M * N * 4)
    tc0 = b.bla_tc_launch_count()
    b.bla_gemm(ta, tb, M, N, K, Ad, M if ta else K, Bd, K if tb else N, Cd, N)
    out = np.empty((M, N), np.float32); b.bla_copy_d2h(ptr(out), Cd, out.nbytes); b.bla_sync()
    err_cols = np.abs(out - want).max(axis=0) / np.abs(want).max()
    print("narrow=%s %s%s %dx%dx%d tc_launches %d rel_err %.3e worst col %d (%.2e)" % (os.environ.get("BLA_TC_NARROW", "1"), "T" if ta else "N", "T" if tb else "N",
          M, N, K, b.bla_tc_launch_count() - tc0, rel_err(out, want), int(np.argmax(err_cols)), err_cols.max()))
    for p_ in (Ad, Bd, Cd):
        b.bla_free(p_)
