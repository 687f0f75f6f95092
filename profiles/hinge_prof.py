"""Capture target: device-resident hinge iterations over 60,000 synthetic MNIST-shaped samples."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0)
n = 60000
rng = np.random.default_rng(3)
x = rng.integers(0, 256, (n, 784)).astype(np.float32); y = rng.integers(0, 10, n).astype(np.float32)
store = b.bla_mnist_from_arrays(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), n, 784)
h = b.bla_hinge_create(784, 10, n)
w0 = (rng.random((10, 784)) / 10 - 0.05).astype(np.float32)
b.bla_hinge_set_weights(h, w0.ctypes.data_as(C.c_void_p))
for _ in range(4):
    b.bla_hinge_iteration(h, store, 0.001, None)
b.bla_sync()
print("done")
