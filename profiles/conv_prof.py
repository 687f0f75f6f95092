"""Capture target: the batched conv2d entry points at the U-Net's shapes, tensor path (for ncu launch lists)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0)
b.bla_set_gemm_path(b.GEMM_3XTF32 if os.environ.get("CONV_PATH", "tc") == "tc" else b.GEMM_FP32)
imgs = int(os.environ.get("CONV_IMGS", "64"))
reps = int(os.environ.get("CONV_REPS", "2"))
for (Cn, H, F, k, st) in ((128, 32, 128, 3, 1), (128, 32, 256, 3, 2), (256, 16, 256, 3, 1), (256, 8, 256, 3, 1), (256, 16, 256, 1, 1)):
    Ho = -(-H // st)
    nx, nw, ny = imgs * Cn * H * H, F * Cn * k * k, imgs * F * Ho * Ho
    x = b.bla_malloc_device(nx * 4); w = b.bla_malloc_device(nw * 4); y = b.bla_malloc_device(ny * 4)
    gx = b.bla_malloc_device(nx * 4); gw = b.bla_malloc_device(nw * 4)
    b.bla_fill_uniform(x, nx, 8, -1, 1); b.bla_fill_uniform(w, nw, 9, -0.05, 0.05); b.bla_fill_uniform(y, ny, 10, -1, 1)
    for _ in range(reps):
        b.bla_conv2d_forward(x, w, y, imgs, Cn, H, H, F, k, st)
        b.bla_conv2d_wgrad(x, y, gw, imgs, Cn, H, H, F, k, st)
        b.bla_conv2d_dgrad(y, w, gx, imgs, Cn, H, H, F, k, st)
    b.bla_sync()
    if os.environ.get("CONV_TIME"):
        flop = 2.0 * F * (imgs * Ho * Ho) * (Cn * k * k)
        res = []
        for name, fn in (("fprop", lambda: b.bla_conv2d_forward(x, w, y, imgs, Cn, H, H, F, k, st)),
                         ("wgrad", lambda: b.bla_conv2d_wgrad(x, y, gw, imgs, Cn, H, H, F, k, st)),
                         ("dgrad", lambda: b.bla_conv2d_dgrad(y, w, gx, imgs, Cn, H, H, F, k, st))):
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            b.bla_sync()
            us = (time.perf_counter() - t0) / 20 * 1e6
            res.append(f"{name} {us:7.1f} us {flop / us / 1e6:6.1f} TF/s")
        print((imgs, Cn, H, F, k, st), " | ".join(res), flush=True)
    for p in (x, w, y, gx, gw):
        b.bla_free(p)
print("done")
