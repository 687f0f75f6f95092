"""Runs every distinct conv shape of the U-Net at a given batch on the tensor path, one op at a time with a sync after each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
b.bla_init(0); b.bla_set_gemm_path(b.GEMM_3XTF32)
shapes = [(128, 32, 128, 3, 1), (128, 32, 256, 3, 2), (256, 16, 256, 3, 1), (512, 16, 256, 3, 1), (512, 16, 256, 1, 1), (256, 16, 256, 3, 2),
          (256, 8, 256, 3, 1), (512, 8, 256, 3, 1), (512, 8, 256, 1, 1), (256, 8, 256, 3, 2), (256, 4, 256, 3, 1), (512, 4, 256, 3, 1),
          (512, 4, 256, 1, 1), (256, 32, 128, 3, 1), (256, 32, 128, 1, 1), (3, 32, 128, 3, 1), (3, 32, 128, 1, 1), (128, 32, 3, 3, 1)]
for (Cn, H, F, k, st) in shapes:
    Ho = -(-H // st)
    nx, nw, ny = imgs * Cn * H * H, F * Cn * k * k, imgs * F * Ho * Ho
    x = b.bla_malloc_device(nx * 4); w = b.bla_malloc_device(nw * 4); y = b.bla_malloc_device(ny * 4)
    gx = b.bla_malloc_device(nx * 4); gw = b.bla_malloc_device(nw * 4)
    b.bla_fill_uniform(x, nx, 8, -1, 1); b.bla_fill_uniform(w, nw, 9, -0.05, 0.05); b.bla_fill_uniform(y, ny, 10, -1, 1)
    b.bla_sync()
    for name, fn in (("fprop", lambda: b.bla_conv2d_forward(x, w, y, imgs, Cn, H, H, F, k, st)),
                     ("wgrad", lambda: b.bla_conv2d_wgrad(x, y, gw, imgs, Cn, H, H, F, k, st)),
                     ("dgrad", lambda: b.bla_conv2d_dgrad(y, w, gx, imgs, Cn, H, H, F, k, st))):
        print((Cn, H, F, k, st), name, end=" ", flush=True)
        t0 = b.bla_tc_launch_count()
        fn(); b.bla_sync()
        print("ok tc", b.bla_tc_launch_count() - t0, flush=True)
    for p in (x, w, y, gx, gw):
        b.bla_free(p)
print("all ok")
