"""One conv shape on the tensor path, a few launches of fprop / wgrad / dgrad -- target of `ncu --set full` captures.
Usage: python profiles/conv_one.py imgs C H F k stride"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
imgs, Cn, H, F, k, st = [int(v) for v in sys.argv[1:7]]
b.bla_init(0)
b.bla_set_gemm_path(b.GEMM_3XTF32)
Ho = -(-H // st)
nx, nw, ny = imgs * Cn * H * H, F * Cn * k * k, imgs * F * Ho * Ho
x = b.bla_malloc_device(nx * 4); w = b.bla_malloc_device(nw * 4); y = b.bla_malloc_device(ny * 4)
gx = b.bla_malloc_device(nx * 4); gw = b.bla_malloc_device(nw * 4)
b.bla_fill_uniform(x, nx, 8, -1, 1); b.bla_fill_uniform(w, nw, 9, -0.05, 0.05); b.bla_fill_uniform(y, ny, 10, -1, 1)
for _ in range(3):
    b.bla_conv2d_forward(x, w, y, imgs, Cn, H, H, F, k, st)
    b.bla_conv2d_wgrad(x, y, gw, imgs, Cn, H, H, F, k, st)
    b.bla_conv2d_dgrad(y, w, gx, imgs, Cn, H, H, F, k, st)
b.bla_sync()
print("done")
