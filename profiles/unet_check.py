"""Prints the U-Net parity errors instead of asserting (development probe for tests/test_unet_gpu.py)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import ptr, rel_err
import unet_ref, test_unet_gpu as T
b.bla_init(0)
for cfg, imgs, name in ((T.SMALL, 3, "small"), (T.FULL, 2, "full")):
    for quirk in (0, 1):
        for path in (b.GEMM_FP32, b.GEMM_3XTF32):
            b.bla_set_quirks(quirk); b.bla_set_gemm_path(path)
            net, tensors = T.make_net(b, cfg, imgs)
            n = b.bla_unet_num_params(net)
            flat = unet_ref.synthetic_params(cfg, tensors, imgs, 5); b.bla_unet_set_params(net, ptr(flat))
            x, temb, noise = T.inputs(cfg, imgs, 3)
            t0 = time.time()
            wo, wl, wg = unet_ref.reference_step(cfg, tensors, flat, x, temb, noise, quirk)
            tref = time.time() - t0
            out = np.empty_like(x); b.bla_unet_forward(net, ptr(x), ptr(temb), imgs, ptr(out))
            loss = np.zeros(1); b.bla_unet_train_step(net, ptr(x), ptr(temb), ptr(noise), imgs, 0.0, ptr(loss))
            g = np.empty(n, np.float32); b.bla_unet_get_grads(net, ptr(g))
            errs = sorted(((rel_err(g[o:o + c], wg[o:o + c]), nm) for nm, o, c in tensors), reverse=True)
            print(f"{name} quirk {quirk} path {path}: params {n} fwd err {rel_err(out, wo):.2e} |out| {np.abs(wo).max():.3g} loss {loss[0]:.6g} vs {wl:.6g} "
                  f"worst grads {[(f'{e:.1e}', nm) for e, nm in errs[:3]]} median {errs[len(errs)//2][0]:.1e} (torch ref {tref:.1f}s)", flush=True)
            b.bla_unet_destroy(net)
