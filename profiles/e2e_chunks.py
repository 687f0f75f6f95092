"""Host-batch chunk width sweep for the MLP step (bla_mlp_set_host_chunking): ms per 60,000-column step, float32 and uint8
pinned host batches, wall-clock around synchronous calls (stats_host forces the synchronise).  Not a bench value."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bla_b200 as b
b.bla_init(0)
B = 60000
dims = (C.c_int * 4)(784, 512, 256, 10)
net = b.bla_mlp_create(dims, B)
b.bla_mlp_init_params(net, 1)
b.bla_set_gemm_path(b.GEMM_AUTO)
rng = np.random.default_rng(0)
hx = b.bla_malloc_pinned(784 * B * 4); hx8 = b.bla_malloc_pinned(784 * B); hy = b.bla_malloc_pinned(10 * B * 4)
x8 = rng.integers(0, 256, 784 * B).astype(np.uint8)
np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(784 * B,))[:] = x8
np.ctypeslib.as_array(C.cast(hx8, C.POINTER(C.c_ubyte)), shape=(784 * B,))[:] = x8
y = np.ctypeslib.as_array(C.cast(hy, C.POINTER(C.c_float)), shape=(10, B)); y[:] = 0; y[rng.integers(0, 10, B), np.arange(B)] = 1
stats = np.zeros(2)
sp = stats.ctypes.data_as(C.c_void_p)
out = {}
for name, fn, x in (("f32", b.bla_mlp_train_step, hx), ("u8", b.bla_mlp_train_step_u8, hx8)):
    for cols in (0, 2048, 3072, 4096, 6144, 8192, 12288, 15360, 30720):
        b.bla_mlp_set_host_chunking(net, cols)
        for _ in range(3): fn(net, x, hy, B, B, 0, 1e-4, sp)
        t = time.perf_counter()
        for _ in range(10): fn(net, x, hy, B, B, 0, 1e-4, sp)
        out[f"{name}_{cols}"] = round((time.perf_counter() - t) * 100, 4)
print(json.dumps(out))
