"""Prints parity errors of bla_attention_forward/backward against float64 torch autograd (development probe)."""
import os, sys, math
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bla_b200 as b
from helpers import ptr, rel_err
b.bla_init(0)
b.bla_set_gemm_path(b.GEMM_FP32)
def dev(a):
    a = np.ascontiguousarray(a, np.float32); d = b.bla_malloc_device(a.nbytes); b.bla_copy_h2d(d, ptr(a), a.nbytes); b.bla_sync(); return d
def host(d, shape):
    o = np.empty(shape, np.float32); b.bla_copy_d2h(ptr(o), d, o.nbytes); b.bla_sync(); return o
for imgs, Cn, S in ((3, 64, 64), (2, 256, 256), (5, 32, 16), (2, 48, 4)):
    rng = np.random.default_rng(S)
    x = rng.normal(size=(imgs, Cn, S)); wqkv = rng.normal(0, 0.2, (Cn, 48)); wo = rng.normal(0, 0.3, (16, Cn)); bo = rng.normal(size=Cn)
    dout = rng.normal(size=(imgs, Cn, S))
    t = [torch.tensor(a, requires_grad=True) for a in (x, wqkv, wo, bo)]
    z = t[0].transpose(1, 2); qkv = z @ t[1]
    q, k, v = qkv[..., :16], qkv[..., 16:32], qkv[..., 32:]
    p = torch.softmax(q @ k.transpose(1, 2) / 4.0, dim=-1)
    out = ((p @ v) @ t[2] + t[3]).transpose(1, 2)
    (out * torch.tensor(dout)).sum().backward()
    dx_, dwqkv_, dwo_, dbo_ = [a.grad.numpy() for a in t]
    M = imgs * S
    xd, wq, wod, bod, dd = dev(x), dev(wqkv), dev(wo), dev(bo), dev(dout)
    zb, qb, pb, ab, ob = [b.bla_malloc_device(n * 4) for n in (M * Cn, M * 48, M * S, M * 16, M * Cn)]
    gq, gw, gb, gx = [b.bla_malloc_device(n * 4) for n in (Cn * 48, 16 * Cn, Cn, M * Cn)]
    b.bla_attention_forward(xd, wq, wod, bod, zb, qb, pb, ab, ob, imgs, Cn, S)
    b.bla_attention_backward(dd, wq, wod, zb, qb, pb, ab, gq, gw, gb, gx, imgs, Cn, S)
    gqh = host(gq, (Cn, 48))
    print((imgs, Cn, S), "out %.1e probs %.1e | dwo %.1e dbo %.1e dQ %.1e dK %.1e dV %.1e dx %.1e" % (
        rel_err(host(ob, out.shape), out.detach().numpy()), rel_err(host(pb, p.shape), p.detach().numpy()),
        rel_err(host(gw, wo.shape), dwo_), rel_err(host(gb, bo.shape), dbo_),
        rel_err(gqh[:, :16], dwqkv_[:, :16]), rel_err(gqh[:, 16:32], dwqkv_[:, 16:32]), rel_err(gqh[:, 32:], dwqkv_[:, 32:]),
        rel_err(host(gx, x.shape), dx_)), flush=True)
