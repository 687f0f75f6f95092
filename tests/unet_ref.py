"""Test-only float64 restatement of model/cifar_unet.c's forward() (:999-1168) in torch, batched, used to check the
device U-Net of csrc/unet.cu: its forward against the oracle's operators (pinned op by op in test_unet_cpu.py) and its
backward against torch autograd (the reference's own backward() is work in progress).  Never imported by the
product."""
import math

import numpy as np
import torch
import torch.nn.functional as F

KEY_DIM = 16


def same_pad(n, k, s):
    """lib/conv.c:12-24: TensorFlow 'SAME' padding, smaller half first"""
    no = -(-n // s)
    p = max(0, (no - 1) * s + k - n)
    return p // 2, p - p // 2


def conv(x, w, stride):
    k = w.shape[-1]
    pt, pb = same_pad(x.shape[2], k, stride)
    pl, pr = same_pad(x.shape[3], k, stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, stride=stride)


def group_norm(x, group_size, quirk):
    """lib/norm.c:5-50: groups of `group_size` channels (the last one may be short); quirk = divide by the variance"""
    outs = []
    for c0 in range(0, x.shape[1], group_size):
        g = x[:, c0:c0 + group_size]
        mu = g.mean(dim=(1, 2, 3), keepdim=True)
        var = ((g - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
        outs.append((g - mu) / (var if quirk else torch.sqrt(var + 1e-8)))
    return torch.cat(outs, dim=1)


class RefUnet:
    """Walks the same node list as bla_unet_create (node ids matter for the dropout seeds)."""

    def __init__(self, cfg, tensors, flat, quirk, drop_masks=None):
        self.cfg, self.quirk = cfg, quirk
        self.p = {}
        for name, off, n in tensors:
            self.p[name] = flat[off:off + n]
        self.node = 0
        self.drop_masks = drop_masks or {}

    def _next(self):
        self.node += 1
        return self.node

    def res(self, name, x, temb, cout):
        nid = self._next()
        c = self.cfg
        cin, k = x.shape[1], c["kernel_size"]
        P = self.p
        h = torch.relu(group_norm(x, c["group_size"], self.quirk))
        h = conv(h, P[name + "/conv_1"].view(cout, cin, k, k), 1)
        td = temb @ P[name + "/time_weight"].view(c["time_dim"], cout) + P[name + "/time_bias"]
        h = h + td[:, :, None, None]
        h = torch.relu(group_norm(h, c["group_size"], self.quirk))
        if nid in self.drop_masks:
            h = h * self.drop_masks[nid].view_as(h)
        o = conv(h, P[name + "/conv_2"].view(cout, cout, k, k), 1)
        r = conv(x, P[name + "/residual_conv"].view(cout, cin, 1, 1), 1) if cin != cout else x
        return o + r

    def attn(self, name, x):
        self._next()
        N, Cn, H, W = x.shape
        P = self.p
        z = x.flatten(2).transpose(1, 2)                                  # (S, C) per image
        qkv = z @ P[name + "/qkv"].view(Cn, 3 * KEY_DIM)
        q, k, v = qkv[..., :KEY_DIM], qkv[..., KEY_DIM:2 * KEY_DIM], qkv[..., 2 * KEY_DIM:]
        s = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(KEY_DIM), dim=-1)
        d = (s @ v) @ P[name + "/weight"].view(KEY_DIM, Cn) + P[name + "/bias"]
        return d.transpose(1, 2).reshape(N, Cn, H, W)

    def convl(self, name, x, cout, k, stride):
        self._next()
        return conv(x, self.p[name].view(cout, x.shape[1], k, k), stride)

    def up(self, x):
        self._next()
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    def cat(self, a, b):
        self._next()
        return torch.cat([a, b], dim=1)

    def forward(self, x, temb):
        c = self.cfg
        D, K = c["dims"], c["kernel_size"]
        self.node = 0
        d1r1 = self.res("down_1/resnet_1", x, temb, D[0])
        d1r2 = self.res("down_1/resnet_2", d1r1, temb, D[0])
        d1c = self.convl("down_1/conv", d1r2, D[1], K, 2)
        d2r1 = self.res("down_2/resnet_1", d1c, temb, D[1])
        d2a1 = self.attn("down_2/self_attention_1", d2r1)
        d2r2 = self.res("down_2/resnet_2", d2a1, temb, D[1])
        d2a2 = self.attn("down_2/self_attention_2", d2r2)
        d2c = self.convl("down_2/conv", d2a2, D[2], K, 2)
        d3r1 = self.res("down_3/resnet_1", d2c, temb, D[2])
        d3r2 = self.res("down_3/resnet_2", d3r1, temb, D[2])
        d3c = self.convl("down_3/conv", d3r2, D[3], K, 2)
        d4r1 = self.res("down_4/resnet_1", d3c, temb, D[3])
        d4r2 = self.res("down_4/resnet_2", d4r1, temb, D[3])
        m1 = self.res("mid/resnet_1", d4r2, temb, D[3])
        ma = self.attn("mid/self_attention", m1)
        m2 = self.res("mid/resnet_2", ma, temb, D[3])
        h = self.cat(m2, d4r2)
        h = self.res("up_1/resnet_1", h, temb, D[3])
        h = self.res("up_1/resnet_2", h, temb, D[3])
        h = self.up(h)
        if D[3] != D[2]:
            h = self.convl("up_1/conv", h, D[2], K, 1)
        h = self.cat(h, d3r2)
        h = self.res("up_2/resnet_1", h, temb, D[2])
        h = self.res("up_2/resnet_2", h, temb, D[2])
        h = self.up(h)
        if D[2] != D[1]:
            h = self.convl("up_2/conv", h, D[1], K, 1)
        h = self.cat(h, d2r2)
        h = self.res("up_3/resnet_1", h, temb, D[1])
        h = self.attn("up_3/self_attention_1", h)
        h = self.res("up_3/resnet_2", h, temb, D[1])
        h = self.attn("up_3/self_attention_2", h)
        h = self.up(h)
        if D[1] != D[0]:
            h = self.convl("up_3/conv", h, D[0], K, 1)
        h = self.cat(h, d1r2)
        h = self.res("up_4/resnet_1", h, temb, D[0])
        h = self.res("up_4/resnet_2", h, temb, D[0])
        self._next()
        h = torch.relu(group_norm(h, c["group_size"], self.quirk))
        return self.convl("output_conv", h, 3, K, 1)


def reference_step(cfg, tensors, flat_np, x, temb, noise, quirk, drop_masks=None):
    """-> (out, loss_sum, grads) in float64: loss_sum = sum over images of the per-image MSE (cifar_unet.c:1858-1872);
    grads = d/dparams of sum (out - noise)^2, i.e. the reference's dY = 2 (out - noise) (:1353-1365) pulled back."""
    flat = torch.tensor(np.asarray(flat_np, np.float64), requires_grad=True)
    net = RefUnet(cfg, tensors, flat, quirk, drop_masks)
    xt, tt, nt = (torch.tensor(np.asarray(a, np.float64)) for a in (x, temb, noise))
    out = net.forward(xt, tt)
    sq = ((out - nt) ** 2).sum()
    sq.backward()
    per_image = 3 * cfg["image_side"] ** 2
    return out.detach().numpy(), float(sq) / per_image, flat.grad.numpy()


def synthetic_params(cfg, tensors, imgs, seed):
    """Well-conditioned float32 parameters for the parity tests: U(-a, a) with a = sqrt(3 / true fan-in) (unit gain), non-zero
    biases, and Q/K projections scaled down so the softmax is not saturated.  (init_parameters' own fan-ins -- the level's pixel
    count, cifar_unet.c:1804-1851 -- blow the attention logits up to one-hot rows, where the float32 gradient of ANY
    implementation is dominated by cancellation.)"""
    fans = {}

    class Probe(RefUnet):
        def res(self, name, x, temb, cout):
            self._next()
            k2 = self.cfg["kernel_size"] ** 2
            fans[name + "/conv_1"] = x.shape[1] * k2
            fans[name + "/conv_2"] = cout * k2
            fans[name + "/residual_conv"] = x.shape[1]
            fans[name + "/time_weight"] = self.cfg["time_dim"]
            return torch.zeros(x.shape[0], cout, x.shape[2], x.shape[3])

        def attn(self, name, x):
            self._next()
            fans[name + "/qkv"] = 4 * x.shape[1]
            fans[name + "/weight"] = KEY_DIM
            return x

        def convl(self, name, x, cout, k, stride):
            self._next()
            fans[name] = x.shape[1] * k * k
            return torch.zeros(x.shape[0], cout, -(-x.shape[2] // stride), -(-x.shape[3] // stride))

    Probe(cfg, tensors, torch.zeros(1), 0).forward(torch.zeros(imgs, 3, cfg["image_side"], cfg["image_side"]), torch.zeros(imgs, cfg["time_dim"]))
    rng = np.random.default_rng(seed)
    total = max(off + n for _, off, n in tensors)
    total = (total + 3) // 4 * 4
    flat = np.zeros(total, np.float32)
    for name, off, n in tensors:
        a = np.sqrt(3.0 / fans[name]) if name in fans else 0.1
        flat[off:off + n] = rng.uniform(-a, a, n)
    return flat
