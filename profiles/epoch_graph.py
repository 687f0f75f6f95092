"""Small-batch MNIST epochs (bla_mlp_train_epoch) with and without the replayed step graph (BLA_MLP_GRAPH): seconds per epoch and
a digest of the trained parameters -- the two must be bit-identical (same kernels, same order).  Run once per setting."""
import ctypes as C, hashlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bla_b200 as b
b.bla_init(0)
libc = C.CDLL(None)
n = 60000
rng = np.random.default_rng(3)
x = rng.integers(0, 256, (n, 784)).astype(np.float32); y = rng.integers(0, 10, n).astype(np.float32)
store = b.bla_mnist_from_arrays(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), n, 784)
out = {"graph": os.environ.get("BLA_MLP_GRAPH", "0")}
for batch in (64, 100, 512):
    dims = (C.c_int * 4)(784, 256, 128, 10)
    net = b.bla_mlp_create(dims, batch)
    b.bla_mlp_init_params(net, 11)
    stats = np.zeros(2)
    sp = stats.ctypes.data_as(C.c_void_p)
    ts = []
    for ep in range(3):
        libc.srand(7 + ep)
        t = time.perf_counter(); b.bla_mlp_train_epoch(net, store, batch, 0.02, sp); ts.append(time.perf_counter() - t)
    shapes = ((256, 784), (256,), (128, 256), (128,), (10, 128), (10,))
    got = [np.empty(s, np.float32) for s in shapes]
    b.bla_mlp_get_params(net, *[g.ctypes.data_as(C.c_void_p) for g in got])
    h = hashlib.sha1(b"".join(g.tobytes() for g in got)).hexdigest()[:16]
    out[f"batch_{batch}"] = {"epoch_s": [round(t, 4) for t in ts], "acc_loss": [float(stats[0]), float(stats[1])], "params_sha1": h}
    b.bla_mlp_destroy(net)
print(json.dumps(out))
