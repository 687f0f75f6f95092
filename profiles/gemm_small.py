"""Times square and MLP-shaped GEMMs on the tensor path (development probe for the tile / split-K model)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0); b.bla_set_gemm_path(b.GEMM_3XTF32)
for (ta, tb, M, N, K) in ((0,0,1024,1024,1024), (0,0,2048,2048,2048), (0,0,512,512,4096), (0,0,768,768,768), (0,1,256,784,60000), (0,1,128,256,60000), (1,0,256,60000,128), (0,0,256,60000,784), (0,0,128,60000,256)):
    A = b.bla_malloc_device(M*K*4); B = b.bla_malloc_device(K*N*4); Cm = b.bla_malloc_device(M*N*4)
    b.bla_fill_uniform(A, M*K, 1, -0.5, 0.5); b.bla_fill_uniform(B, K*N, 2, -0.5, 0.5)
    f = lambda: b.bla_gemm(ta, tb, M, N, K, A, M if ta else K, B, K if tb else N, Cm, N)
    for _ in range(3): f()
    b.bla_sync(); t0 = time.perf_counter()
    for _ in range(20): f()
    b.bla_sync(); us = (time.perf_counter() - t0) / 20 * 1e6
    print((ta, tb, M, N, K), f"{us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s", flush=True)
    for p in (A, B, Cm): b.bla_free(p)
