"""Correctness probe of the tensor path on a few shapes (prints the error instead of asserting)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import ptr, rel_err
b.bla_init(0); b.bla_set_gemm_path(b.GEMM_3XTF32)
rng = np.random.default_rng(0)
for (ta, tb, M, N, K) in [(0, 1, 128, 256, 64), (0, 1, 256, 256, 64), (0, 0, 256, 256, 64), (1, 0, 256, 512, 96), (1, 1, 512, 300, 128), (0, 0, 152, 204, 76),
                          (0, 1, 256, 784, 4096), (0, 0, 1024, 2048, 512)]:
    a = rng.uniform(-0.5, 0.5, (M, K)); bm = rng.uniform(-0.5, 0.5, (K, N))
    a_st = np.ascontiguousarray(a.T if ta else a, np.float32); b_st = np.ascontiguousarray(bm.T if tb else bm, np.float32)
    want = a_st.astype(np.float64).T @ (b_st.astype(np.float64).T if tb else b_st) if ta else a_st.astype(np.float64) @ (b_st.astype(np.float64).T if tb else b_st)
    out = np.full((M, N), np.nan, np.float32)
    n0 = b.bla_tc_launch_count()
    b.bla_gemm(ta, tb, M, N, K, ptr(a_st), a_st.shape[1], ptr(b_st), b_st.shape[1], ptr(out), N)
    print((ta, tb, M, N, K), "tc" if b.bla_tc_launch_count() > n0 else "simt", "err %.3e" % rel_err(out, want), "nan" if np.isnan(out).any() else "", flush=True)
