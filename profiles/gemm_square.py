"""Times the square GEMM sweep on the tensor path (development probe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0); b.bla_set_gemm_path(b.GEMM_3XTF32)
for n in (4096, 8192, 16384):
    A = b.bla_malloc_device(n*n*4); B = b.bla_malloc_device(n*n*4); Cm = b.bla_malloc_device(n*n*4)
    b.bla_fill_uniform(A, n*n, 1, -0.5, 0.5); b.bla_fill_uniform(B, n*n, 2, -0.5, 0.5)
    f = lambda: b.bla_gemm(0, 0, n, n, n, A, n, B, n, Cm, n)
    for _ in range(2): f()
    b.bla_sync(); it = 10 if n <= 8192 else 3; t0 = time.perf_counter()
    for _ in range(it): f()
    b.bla_sync(); ms = (time.perf_counter() - t0) / it * 1e3
    print(n, f"{ms:8.3f} ms {2.0*n**3/ms/1e9:7.1f} TF/s", flush=True)
    for p in (A, B, Cm): b.bla_free(p)
