import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import ptr, rel_err
import unet_ref, test_unet_gpu as T
b.bla_init(0)
cfg, imgs = T.SMALL, 3
b.bla_set_quirks(0); b.bla_set_gemm_path(b.GEMM_FP32)
net, tensors = T.make_net(b, cfg, imgs)
b.bla_unet_init_params(net, 5)
n = b.bla_unet_num_params(net)
flat = np.empty(n, np.float32); b.bla_unet_get_params(net, ptr(flat))
x, temb, noise = T.inputs(cfg, imgs, 3)
wo, wl, wg = unet_ref.reference_step(cfg, tensors, flat, x, temb, noise, 0)
loss = np.zeros(1); b.bla_unet_train_step(net, ptr(x), ptr(temb), ptr(noise), imgs, 0.0, ptr(loss))
g = np.empty(n, np.float32); b.bla_unet_get_grads(net, ptr(g))
for nm, o, c in tensors:
    print(f"{nm:40s} err {rel_err(g[o:o+c], wg[o:o+c]):.2e}  |got| {np.abs(g[o:o+c]).max():.3e} |want| {np.abs(wg[o:o+c]).max():.3e}")
