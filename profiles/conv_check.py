"""Probe of the tensor-path implicit-GEMM conv (prints errors instead of asserting)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import load_oracle, ptr, rel_err
b.bla_init(0)
o = load_oracle(np.float64)
for (imgs, Cin, H, W, F, k, s) in [(3, 128, 32, 32, 128, 3, 1), (4, 256, 16, 16, 256, 3, 1), (2, 128, 32, 32, 256, 3, 2), (5, 256, 8, 8, 256, 1, 1), (8, 256, 4, 4, 256, 3, 1), (9, 128, 16, 16, 128, 3, 2), (3, 256, 16, 16, 256, 1, 2), (6, 64, 8, 8, 128, 3, 1), (64, 256, 8, 8, 256, 3, 2), (48, 512, 4, 4, 256, 3, 1), (32, 512, 4, 4, 256, 1, 1), (64, 64, 4, 4, 128, 3, 2), (4, 3, 32, 32, 128, 3, 1), (4, 128, 32, 32, 3, 3, 1), (4, 3, 32, 32, 128, 1, 1), (5, 40, 16, 16, 72, 3, 1)]:
    rng = np.random.default_rng(1)
    Ho, Wo = -(-H // s), -(-W // s)
    x = rng.normal(size=(imgs, Cin, H, W)); kr = rng.normal(0, 0.05, (F, Cin, k, k)); dy = rng.normal(size=(imgs, F, Ho, Wo))
    y = np.empty((imgs, F, Ho, Wo)); dx = np.empty_like(x); dk = np.zeros_like(kr)
    for n in range(imgs):
        o.orc_conv(Cin, H, W, F, k, s, ptr(x[n]), ptr(kr), ptr(y[n]))
        dkn = np.empty_like(kr)
        o.orc_conv_ddx(Cin, H, W, F, k, s, ptr(x[n]), ptr(kr), ptr(dy[n]), ptr(dkn), ptr(dx[n])); dk += dkn
    def dev(a):
        a = np.ascontiguousarray(a, np.float32); d = b.bla_malloc_device(a.nbytes); b.bla_copy_h2d(d, ptr(a), a.nbytes); return d
    xd, wd, dyd = dev(x), dev(kr), dev(dy)
    dwd = b.bla_malloc_device(kr.size * 4)
    yd = b.bla_malloc_device(y.size * 4); dxd = b.bla_malloc_device(x.size * 4)
    for path in (b.GEMM_FP32, b.GEMM_3XTF32):
        b.bla_set_gemm_path(path)
        n0 = b.bla_tc_launch_count()
        b.bla_conv2d_forward(xd, wd, yd, imgs, Cin, H, W, F, k, s)
        b.bla_conv2d_dgrad(dyd, wd, dxd, imgs, Cin, H, W, F, k, s)
        b.bla_conv2d_wgrad(xd, dyd, dwd, imgs, Cin, H, W, F, k, s)
        odw = np.empty(kr.shape, np.float32); b.bla_copy_d2h(ptr(odw), dwd, odw.nbytes)
        out = np.empty(y.shape, np.float32); b.bla_copy_d2h(ptr(out), yd, out.nbytes)
        odx = np.empty(x.shape, np.float32); b.bla_copy_d2h(ptr(odx), dxd, odx.nbytes); b.bla_sync()
        print((imgs, Cin, H, W, F, k, s), "path", path, "tc launches", b.bla_tc_launch_count() - n0, "fprop err %.2e dgrad err %.2e wgrad err %.2e" % (rel_err(out, y), rel_err(odx, dx), rel_err(odw, dk)), flush=True)
