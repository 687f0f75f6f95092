import ctypes as C, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bla_b200 as b
b.bla_init(0)
n=60000
rng=np.random.default_rng(3)
x=rng.integers(0,256,(n,784)).astype(np.float32); y=rng.integers(0,10,n).astype(np.float32)
store=b.bla_mnist_from_arrays(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), n, 784)
hg=b.bla_hinge_create(784,10,n)
w0=(rng.random((10,784))/10-0.05).astype(np.float32)
b.bla_hinge_set_weights(hg, w0.ctypes.data_as(C.c_void_p))
for _ in range(3): b.bla_hinge_iteration(hg, store, 0.001, None)
b.bla_sync()
t0=time.perf_counter()
for _ in range(50): b.bla_hinge_iteration(hg, store, 0.001, None)
b.bla_sync()
dt=(time.perf_counter()-t0)/50
print("hinge iteration %.1f us, %.2f TB/s algorithmic" % (dt*1e6, n*784*4/dt/1e12))
