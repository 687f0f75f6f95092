"""Capture / timing target: the batched U-Net training step at the reference's size (cifar_unet.c:26-37).
Usage: python profiles/unet_prof.py [imgs] [steps] [path: tc|fp32]   (UNET_TIME=1 prints ms/step)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import ptr
imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
path = sys.argv[3] if len(sys.argv) > 3 else "tc"
b.bla_init(0)
b.bla_set_gemm_path({"tc": b.GEMM_AUTO, "tc_forced": b.GEMM_3XTF32, "fp32": b.GEMM_FP32}[path])
b.bla_set_quirks(int(os.environ.get("BLA_QUIRKS", "1")))
uc = b.UnetConfig(32, (C.c_int * 4)(128, 256, 256, 256), 512, 3, 32, 16, 0.1, imgs, 7)
net = b.bla_unet_create(C.byref(uc))
b.bla_unet_init_params(net, 42)
n3 = imgs * 3 * 32 * 32
x = b.bla_malloc_device(n3 * 4); nz = b.bla_malloc_device(n3 * 4); te = b.bla_malloc_device(imgs * 512 * 4)
b.bla_fill_uniform(x, n3, 1, -1, 1); b.bla_fill_uniform(nz, n3, 2, -1, 1); b.bla_fill_uniform(te, imgs * 512, 3, -1, 1)
loss = np.zeros(1)
l0 = b.bla_launch_count()
b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, ptr(loss))
print("launches per step", b.bla_launch_count() - l0, "loss_sum", loss[0], "params", b.bla_unet_num_params(net), flush=True)
for _ in range(steps - 1):
    b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, None)
b.bla_sync()
if os.environ.get("UNET_TIME"):
    t0 = time.perf_counter()
    for _ in range(5):
        b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, None)
    b.bla_sync()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    out = b.bla_malloc_device(n3 * 4)
    t0 = time.perf_counter()
    for _ in range(5):
        b.bla_unet_forward(net, x, te, imgs, out)
    b.bla_sync()
    fms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"imgs {imgs} path {path}: train step {ms:.2f} ms = {imgs / ms * 1e3:.0f} images/s ({3 * 7.13e9 * imgs / ms / 1e9:.1f} TFLOP/s at 3 x 7.13 GFLOP/image); "
          f"forward {fms:.2f} ms = {imgs / fms * 1e3:.0f} images/s ({7.13e9 * imgs / fms / 1e9:.1f} TFLOP/s)", flush=True)
    b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, ptr(loss)); print("loss after", loss[0])
print("done")
