import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bla_b200 as b
from helpers import ptr, rel_err
b.bla_init(0)
b.bla_set_gemm_path(b.GEMM_3XTF32)
rng = np.random.default_rng(0)
def dev(a):
    d = b.bla_malloc_device(a.nbytes); b.bla_copy_h2d(d, ptr(a), a.nbytes); b.bla_sync(); return d
for (M, N, K) in ((256, 256, 784), (256, 1024, 784), (256, 3000, 784), (256, 7500, 784), (128, 952, 784)):
    for kind in ("int", "rand"):
        A = rng.uniform(-0.08, 0.08, (M, K)).astype(np.float32)
        Bm = (rng.integers(0, 256, (K, N)).astype(np.float32) if kind == "int" else rng.uniform(-0.5, 0.5, (K, N)).astype(np.float32))
        want = A.astype(np.float64) @ Bm.astype(np.float64)
        Ad, Bd = dev(A), dev(Bm); Cd = b.bla_malloc_device(M * N * 4)
        b.bla_gemm(0, 0, M, N, K, Ad, K, Bd, N, Cd, N)
        out = np.empty((M, N), np.float32); b.bla_copy_d2h(ptr(out), Cd, out.nbytes); b.bla_sync()
        print(os.environ.get("BLA_TC_NARROW", "1"), M, N, K, kind, "rel_err %.3e" % rel_err(out, want), "max col err", np.argmax(np.abs(out - want).max(axis=0)))
