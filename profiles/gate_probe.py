"""Times the MLP's layer-2 dgrad GEMM (256 x 60000 x 128, A stored transposed) with and without the relu' gate epilogue."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0); b.bla_set_gemm_path(b.GEMM_3XTF32)
M, N, K = 256, 60000, 128
A = b.bla_malloc_device(M*K*4); B = b.bla_malloc_device(K*N*4); Cm = b.bla_malloc_device(M*N*4); G = b.bla_malloc_device(M*N*4)
b.bla_fill_uniform(A, M*K, 1, -0.5, 0.5); b.bla_fill_uniform(B, K*N, 2, -0.5, 0.5); b.bla_fill_uniform(G, M*N, 3, -1, 1)
for name, gate in (("plain", None), ("gate", G)):
    epi = b.Epilogue(); epi.gate = gate
    f = lambda: b.bla_gemm_ex(1, 0, M, N, K, A, M, B, N, Cm, N, C.byref(epi))
    for _ in range(3): f()
    b.bla_sync(); t0 = time.perf_counter()
    for _ in range(50): f()
    b.bla_sync(); us = (time.perf_counter() - t0) / 50 * 1e6
    print(name, f"{us:7.1f} us {2.0*M*N*K/us/1e6:6.1f} TF/s  ({(M*N*(8 if gate else 4) + K*N*4)/us/1e3:.0f} GB/s of algorithmic bytes)", flush=True)
