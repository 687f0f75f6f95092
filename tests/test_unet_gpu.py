"""GPU parity of the batched device-resident U-Net step (csrc/unet.cu, include/bla.h "CIFAR U-Net trainer") against the
float64 checker of tests/unet_ref.py (pinned to the oracle in test_unet_cpu.py): forward() of model/cifar_unet.c:1099-1168
under the reference's semantics (BLA_QUIRKS=1) and the intended ones, and the gradients of every parameter tensor against
autograd (the reference's own backward() is work in progress)."""
import ctypes as C

import numpy as np
import pytest

from helpers import ptr, rel_err
import unet_ref

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def bla():
    import bla_b200 as b
    assert b.bla_device_count() >= 1
    assert not b.MISSING
    b.bla_set_gemm_path(b.GEMM_FP32)
    b.bla_set_quirks(1)
    return b


SMALL = dict(image_side=16, dims=(32, 64, 64, 32), time_dim=24, kernel_size=3, group_size=8, key_dim=16)
FULL = dict(image_side=32, dims=(128, 256, 256, 256), time_dim=512, kernel_size=3, group_size=32, key_dim=16)   # cifar_unet.c:26-37


def make_net(b, cfg, max_imgs, dropout=0.0, seed=11):
    uc = b.UnetConfig(cfg["image_side"], (C.c_int * 4)(*cfg["dims"]), cfg["time_dim"], cfg["kernel_size"], cfg["group_size"],
                      cfg["key_dim"], dropout, max_imgs, seed)
    net = b.bla_unet_create(C.byref(uc))
    tensors = [(b.bla_unet_tensor_name(net, i).decode(), b.bla_unet_tensor_offset(net, i), b.bla_unet_tensor_size(net, i))
               for i in range(b.bla_unet_num_tensors(net))]
    return net, tensors


def inputs(cfg, imgs, seed):
    rng = np.random.default_rng(seed)
    side, T = cfg["image_side"], cfg["time_dim"]
    x = rng.uniform(-1, 1, (imgs, 3, side, side)).astype(np.float32)
    # sinusoidal time embedding of a random diffusion step per image (the reference leaves its own uninitialised, D6)
    t = rng.integers(0, 1000, imgs)[:, None]
    freq = np.exp(-np.log(10000.0) * np.arange(T // 2) / (T // 2))[None]
    temb = np.concatenate([np.sin(t * freq), np.cos(t * freq)], axis=1).astype(np.float32)
    noise = rng.normal(size=x.shape).astype(np.float32)
    return x, temb, noise


def dropout_masks(b, cfg, tensors, imgs, rate, seed, step):
    """keep-masks per ResNet node id, from the documented generator (include/bla.h bla_unet_config.seed)"""
    import torch
    sizes = {}

    class Probe(unet_ref.RefUnet):
        def res(self, name, x, temb, cout):
            sizes[self._next()] = (x.shape[0], cout, x.shape[2], x.shape[3])
            return torch.zeros(x.shape[0], cout, x.shape[2], x.shape[3])

        def attn(self, name, x):
            self._next()
            return x

        def convl(self, name, x, cout, k, stride):
            self._next()
            return torch.zeros(x.shape[0], cout, -(-x.shape[2] // stride), -(-x.shape[3] // stride))

    Probe(cfg, tensors, torch.zeros(1), 0).forward(torch.zeros(imgs, 3, cfg["image_side"], cfg["image_side"]),
                                                    torch.zeros(imgs, cfg["time_dim"]))
    masks = {}
    for nid, shape in sizes.items():
        cnt = int(np.prod(shape))
        u = np.empty(cnt, np.float32)
        b.bla_host_uniform(ptr(u), cnt, seed + 7919 * step + nid, 0.0, 1.0)
        masks[nid] = torch.tensor((u >= np.float32(rate)).astype(np.float64).reshape(shape))
    return masks


def run_case(b, cfg, imgs, quirk, path, check_grads, fwd_tol, grad_tol, dropout=0.0, own_init=False):
    b.bla_set_quirks(quirk)
    b.bla_set_gemm_path(path)
    net, tensors = make_net(b, cfg, imgs, dropout)
    try:
        n = b.bla_unet_num_params(net)
        if own_init:
            b.bla_unet_init_params(net, 5)
            flat = np.empty(n, np.float32)
            b.bla_unet_get_params(net, ptr(flat))
        else:
            flat = unet_ref.synthetic_params(cfg, tensors, imgs, 5)
            assert flat.size == n
            b.bla_unet_set_params(net, ptr(flat))
        x, temb, noise = inputs(cfg, imgs, 3)
        masks = dropout_masks(b, cfg, tensors, imgs, dropout, 11, step=0) if dropout > 0 else None
        want_out, want_loss, want_g = unet_ref.reference_step(cfg, tensors, flat, x, temb, noise, quirk, masks)
        if dropout == 0:
            out = np.empty_like(x)
            b.bla_unet_forward(net, ptr(x), ptr(temb), imgs, ptr(out))
            assert np.isfinite(out).all()
            assert rel_err(out, want_out) <= fwd_tol, ("forward", rel_err(out, want_out))
        if not check_grads:
            return
        loss = np.zeros(1)
        b.bla_unet_train_step(net, ptr(x), ptr(temb), ptr(noise), imgs, 0.0, ptr(loss))
        assert abs(loss[0] - want_loss) <= max(fwd_tol, 1e-4) * abs(want_loss), (loss[0], want_loss)
        g = np.empty(n, np.float32)
        b.bla_unet_get_grads(net, ptr(g))
        worst = ("", 0.0)
        for name, off, cnt in tensors:
            e = rel_err(g[off:off + cnt], want_g[off:off + cnt])
            if e > worst[1]:
                worst = (name, e)
        assert worst[1] <= grad_tol, ("gradient", worst)
        if dropout == 0:
            # SGD: params -= lr * grads (the optimiser the reference never got to, cifar_unet.c:1887-1888)
            b.bla_unet_train_step(net, ptr(x), ptr(temb), ptr(noise), imgs, 0.5, None)
            after = np.empty(n, np.float32)
            b.bla_unet_get_params(net, ptr(after))
            g2 = np.empty(n, np.float32)
            b.bla_unet_get_grads(net, ptr(g2))
            assert rel_err(after, flat - np.float32(0.5) * g2) <= 1e-6
    finally:
        b.bla_unet_destroy(net)
        b.bla_set_quirks(1)
        b.bla_set_gemm_path(b.GEMM_FP32)


@pytest.mark.parametrize("quirk", [0, 1])
def test_unet_small_fp32_vs_autograd(bla, quirk):
    # quirk 1 = the reference's forward (group norm divides by the variance); its group_norm_ddx is then NOT the adjoint of
    # that forward (lib/norm.c:52-93 is the textbook formula with the variance in place of sigma), so gradients are compared
    # with autograd under the intended semantics only
    run_case(bla, SMALL, 3, quirk, bla.GEMM_FP32, check_grads=(quirk == 0), fwd_tol=1e-5, grad_tol=1e-5)


def test_unet_small_3xtf32_vs_autograd(bla):
    run_case(bla, SMALL, 3, 0, bla.GEMM_3XTF32, check_grads=True, fwd_tol=1e-4, grad_tol=1e-4)


def test_unet_small_dropout_mask_and_gradients(bla):
    run_case(bla, SMALL, 2, 0, bla.GEMM_FP32, check_grads=True, fwd_tol=1e-5, grad_tol=1e-5, dropout=0.1)


def test_unet_forward_with_reference_init(bla):
    """init_parameters' He / Xavier uniform with the reference's fan-ins (cifar_unet.c:1804-1851): forward only (see
    unet_ref.synthetic_params for why gradients are compared on better-conditioned parameters)."""
    run_case(bla, SMALL, 2, 1, bla.GEMM_FP32, check_grads=False, fwd_tol=2e-5, grad_tol=0, own_init=True)


@pytest.mark.parametrize("path", ["fp32", "3xtf32"])
def test_unet_reference_size_vs_autograd(bla, path):
    """The reference's own configuration (cifar_unet.c:26-37), 2 images: forward and every parameter gradient."""
    p = bla.GEMM_FP32 if path == "fp32" else bla.GEMM_3XTF32
    # FP32 path: <= 1e-5 (north star); 3xTF32: <= 1e-3 per GEMM (north star), 2e-3 through the ~140 chained GEMMs of the gradient
    run_case(bla, FULL, 2, 0, p, check_grads=True, fwd_tol=1e-5 if path == "fp32" else 1e-4, grad_tol=1e-5 if path == "fp32" else 2e-3)


def _dev(b, a):
    a = np.ascontiguousarray(a, np.float32)
    d = b.bla_malloc_device(a.nbytes)
    b.bla_copy_h2d(d, ptr(a), a.nbytes)
    b.bla_sync()
    return d


def _host(b, d, shape):
    o = np.empty(shape, np.float32)
    b.bla_copy_d2h(ptr(o), d, o.nbytes)
    b.bla_sync()
    return o


@pytest.mark.parametrize("imgs,Cn,S", [(3, 64, 64), (2, 256, 256), (5, 32, 16), (2, 48, 4), (1, 40, 100)])
def test_fused_attention_vs_autograd(bla, imgs, Cn, S):
    """_forward_attention / _backward_attention (cifar_unet.c:999-1022, :1261-1335) as the fused device block; S = 256 and 16 are
    the reference's two sizes.  Checker: the same composition in float64 (softmax pinned to lib/util.c in test_unet_cpu.py)."""
    import torch
    b = bla
    b.bla_set_gemm_path(b.GEMM_FP32)
    rng = np.random.default_rng(S + Cn)
    x = rng.normal(size=(imgs, Cn, S)); wqkv = rng.normal(0, 0.2, (Cn, 48)); wo = rng.normal(0, 0.3, (16, Cn)); bo = rng.normal(size=Cn)
    dout = rng.normal(size=(imgs, Cn, S))
    t = [torch.tensor(a.astype(np.float32).astype(np.float64), requires_grad=True) for a in (x, wqkv, wo, bo)]
    qkv = t[0].transpose(1, 2) @ t[1]
    q, k, v = qkv[..., :16], qkv[..., 16:32], qkv[..., 32:]
    p = torch.softmax(q @ k.transpose(1, 2) / 4.0, dim=-1)
    out = ((p @ v) @ t[2] + t[3]).transpose(1, 2)
    (out * torch.tensor(dout.astype(np.float32).astype(np.float64))).sum().backward()
    M = imgs * S
    xd, wq, wod, bod, dd = (_dev(b, a) for a in (x, wqkv, wo, bo, dout))
    bufs = [b.bla_malloc_device(n * 4) for n in (M * Cn, M * 48, M * S, M * 16, M * Cn, Cn * 48, 16 * Cn, Cn, M * Cn)]
    zb, qb, pb, ab, ob, gq, gw, gb, gx = bufs
    b.bla_attention_forward(xd, wq, wod, bod, zb, qb, pb, ab, ob, imgs, Cn, S)
    b.bla_attention_backward(dd, wq, wod, zb, qb, pb, ab, gq, gw, gb, gx, imgs, Cn, S)
    try:
        assert rel_err(_host(b, ob, out.shape), out.detach().numpy()) <= 1e-5
        assert rel_err(_host(b, pb, p.shape), p.detach().numpy()) <= 1e-5
        for got, want in ((gx, t[0]), (gq, t[1]), (gw, t[2]), (gb, t[3])):
            assert rel_err(_host(b, got, want.shape), want.grad.numpy()) <= 1e-5
    finally:
        for d in [xd, wq, wod, bod, dd] + bufs:
            b.bla_free(d)


def test_unet_single_image_and_host_or_device_inputs(bla):
    """the reference's own batch size (one image per forward) and device-resident inputs give the same result as host inputs"""
    b = bla
    b.bla_set_quirks(1)
    net, tensors = make_net(b, SMALL, 2)
    try:
        b.bla_unet_init_params(net, 3)
        x, temb, noise = inputs(SMALL, 2, 9)
        both = np.empty_like(x)
        b.bla_unet_forward(net, ptr(x), ptr(temb), 2, ptr(both))
        one = np.empty_like(x[:1])
        b.bla_unet_forward(net, ptr(np.ascontiguousarray(x[1:])), ptr(np.ascontiguousarray(temb[1:])), 1, ptr(one))
        assert rel_err(one[0], both[1]) <= 1e-6                    # images are independent: no batch statistics anywhere
        xd, td, od = _dev(b, x), _dev(b, temb), b.bla_malloc_device(x.nbytes)
        h0 = b.bla_h2d_bytes()
        b.bla_unet_forward(net, xd, td, 2, od)
        assert b.bla_h2d_bytes() == h0                             # nothing staged
        assert np.array_equal(_host(b, od, x.shape), both)
        for d in (xd, td, od):
            b.bla_free(d)
    finally:
        b.bla_unet_destroy(net)


def test_unet_sgd_reduces_the_loss_on_a_fixed_batch(bla):
    """the trainer as a trainer: plain SGD on one fixed batch (intended group-norm semantics, dropout on) drives the MSE down"""
    b = bla
    b.bla_set_quirks(0)
    b.bla_set_gemm_path(b.GEMM_AUTO)
    net, tensors = make_net(b, SMALL, 4, dropout=0.1)
    try:
        flat = unet_ref.synthetic_params(SMALL, tensors, 4, 1)
        b.bla_unet_set_params(net, ptr(flat))
        x, temb, noise = inputs(SMALL, 4, 2)
        losses = []
        for _ in range(40):
            loss = np.zeros(1)
            # gradients are SUMS over pixels and images of 2 (out - noise) (cifar_unet.c:1353-1365, no 1/N): a small step
            b.bla_unet_train_step(net, ptr(x), ptr(temb), ptr(noise), 4, 3e-6, ptr(loss))
            losses.append(loss[0] / 4)
        assert np.isfinite(losses).all()
        assert losses[-1] < 0.85 * losses[0] and min(losses) > 0, losses
    finally:
        b.bla_unet_destroy(net)
        b.bla_set_quirks(1)
        b.bla_set_gemm_path(b.GEMM_FP32)


def test_unet_reference_size_gradient_is_additive_over_images(bla):
    """The reference-size model on 8 images, tensor path: the gradient of the batch is the sum of the gradients of its halves
    and the losses add (what the data-parallel all-reduce relies on); no CPU checker is needed at this size."""
    b = bla
    b.bla_set_quirks(0)
    b.bla_set_gemm_path(b.GEMM_AUTO)
    net, tensors = make_net(b, FULL, 8)
    try:
        flat = unet_ref.synthetic_params(FULL, tensors, 8, 5)
        b.bla_unet_set_params(net, ptr(flat))
        x, temb, noise = inputs(FULL, 8, 4)
        n = flat.size

        def grads(lo, hi):
            loss = np.zeros(1)
            b.bla_unet_train_step(net, ptr(np.ascontiguousarray(x[lo:hi])), ptr(np.ascontiguousarray(temb[lo:hi])),
                                  ptr(np.ascontiguousarray(noise[lo:hi])), hi - lo, 0.0, ptr(loss))
            g = np.empty(n, np.float32)
            b.bla_unet_get_grads(net, ptr(g))
            return g.astype(np.float64), loss[0]
        g_all, l_all = grads(0, 8)
        g_a, l_a = grads(0, 4)
        g_b, l_b = grads(4, 8)
        assert abs(l_all - (l_a + l_b)) <= 1e-5 * abs(l_all)
        worst = max(rel_err(g_a[o:o + c] + g_b[o:o + c], g_all[o:o + c]) for _, o, c in tensors)
        assert worst <= 2e-3, worst          # the 3xTF32 budget through ~140 chained GEMMs, as in test_unet_reference_size_vs_autograd
    finally:
        b.bla_unet_destroy(net)
        b.bla_set_quirks(1)
        b.bla_set_gemm_path(b.GEMM_FP32)
