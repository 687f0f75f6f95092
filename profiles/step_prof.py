"""One GPU, one data-parallel SHARD of the MLP step: B of the 60,000 global columns, device-resident, no communicator -- what every
rank of an N-GPU strong-scaling run executes apart from the all-reduce.  Usage: python profiles/step_prof.py B [steps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bla_b200 as b

B = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
b.bla_init(0)
stream = torch.cuda.Stream()
b.bla_set_stream(C.c_void_p(stream.cuda_stream))
dims = (C.c_int * 4)(784, 256, 128, 10)
net = b.bla_mlp_create(dims, B)
b.bla_mlp_init_params(net, 42)
rng = np.random.default_rng(0)
nbuf = max(2, -(-(256 << 20) // (784 * B * 4)))
bufs = []
for i in range(nbuf):
    px = rng.integers(0, 256, 784 * B, dtype=np.uint8)
    pd = b.bla_malloc_device(px.nbytes); xd = b.bla_malloc_device(px.nbytes * 4)
    b.bla_copy_h2d(pd, px.ctypes.data_as(C.c_void_p), px.nbytes); b.bla_u8_to_float(xd, pd, px.size, 1.0); b.bla_sync(); b.bla_free(pd)
    lab = rng.integers(0, 10, B); Y = np.zeros((10, B), np.float32); Y[lab, np.arange(B)] = 1
    yd = b.bla_malloc_device(Y.nbytes); b.bla_copy_h2d(yd, Y.ctypes.data_as(C.c_void_p), Y.nbytes); b.bla_sync()
    bufs.append((xd, yd))
for i in range(2 * nbuf + 2):
    b.bla_mlp_train_step(net, bufs[i % nbuf][0], bufs[i % nbuf][1], B, 60000, 0, 0.02, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = b.bla_launch_count()
e0.record(stream)
for i in range(steps):
    b.bla_mlp_train_step(net, bufs[i % nbuf][0], bufs[i % nbuf][1], B, 60000, 0, 0.02, None)
e1.record(stream)
torch.cuda.synchronize()
print("B %d: %.1f us per step, %d launches per step, graphs %s" % (B, e0.elapsed_time(e1) / steps * 1e3, (b.bla_launch_count() - l0) // steps,
                                                                  os.environ.get("BLA_MLP_STEP_GRAPH", "1")))
