"""World-size-2 gloo tests (CPU): the data-parallel decomposition of the MLP step is exact.
Each rank holds a column shard; with the pinned f64 oracle doing the arithmetic, the all-reduced shard
gradients (including the reference's col_sum quirk, decomposed by dp.quirk_window_segments exactly like
csrc/mlp.cu:bias_grad_kernel) must equal the oracle's full-batch step."""
import ctypes as C
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from helpers import ROOT, load_oracle, ptr

dp = importlib.import_module("big-linear-algebra_b200.dp") if os.path.exists(os.path.join(ROOT, "big-linear-algebra_b200", "libbla.so")) else None


def test_shard_columns_partition():
    for B, W in [(60000, 1), (60000, 8), (60001, 8), (7, 3)]:
        cover = []
        for r in range(W):
            off, cnt = dp.shard_columns(B, W, r)
            cover += list(range(off, off + cnt))
        assert cover == list(range(B))


def test_quirk_windows_reassemble_the_reference_col_sum():
    o = load_oracle(np.float64)
    rng = np.random.default_rng(0)
    for rows, Bg, W in [(10, 64, 2), (128, 512, 4), (256, 600, 3), (10, 7, 2), (33, 1000, 8)]:
        m = rng.normal(size=(rows, Bg))
        want = np.empty((rows, 1)); o.orc_col_sum(rows, Bg, ptr(m), ptr(want), 1)
        got = np.zeros(rows)
        for r in range(W):
            off, cnt = dp.shard_columns(Bg, W, r)
            local = np.ascontiguousarray(m[:, off:off + cnt])
            for i in range(rows):
                for (row, a, b) in dp.quirk_window_segments(i, rows, Bg, off, cnt):
                    got[i] += local[row, a:b].sum()
        np.testing.assert_allclose(got, want.ravel(), rtol=1e-12, atol=1e-12)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        o = load_oracle(np.float64)
        dims = [784, 256, 128, 10]
        rng = np.random.default_rng(7)                      # same stream on every rank
        params = [rng.uniform(-0.08, 0.08, s) for s in ((256, 784), (256,), (128, 256), (128,), (10, 128), (10,))]
        X = rng.integers(0, 256, (784, B)).astype(np.float64)
        labels = rng.integers(0, 10, B)
        Y = np.zeros((10, B)); Y[labels, np.arange(B)] = 1
        off, cnt = dp.shard_columns(B, world, rank)
        Xl, Yl = np.ascontiguousarray(X[:, off:off + cnt]), np.ascontiguousarray(Y[:, off:off + cnt])
        # local forward/backward in numpy float64, mirroring csrc/mlp.cu step()
        W1, b1, W2, b2, W3, b3 = params
        A0 = Xl * np.float64(np.float32(1 / np.float32(255.0)))
        Z1 = W1 @ A0 + b1[:, None]; A1 = np.maximum(Z1, 0)
        Z2 = W2 @ A1 + b2[:, None]; A2 = np.maximum(Z2, 0)
        Z3 = W3 @ A2 + b3[:, None]
        P = np.exp(Z3 - Z3.max(0)); P /= P.sum(0)
        dZ3 = (P - Yl) / 784.0
        dZ2 = (W3.T @ dZ3) * (A2 > 0)
        dZ1 = (W2.T @ dZ2) * (A1 > 0)
        grads = []
        for dz, a_prev in ((dZ1, A0), (dZ2, A1), (dZ3, A2)):
            rows = dz.shape[0]
            db = np.zeros(rows)
            for i in range(rows):
                for (row, a, b) in dp.quirk_window_segments(i, rows, B, off, cnt):
                    db[i] += dz[row, a:b].sum()
            grads += [dz @ a_prev.T, db]
        flat = torch.from_numpy(np.concatenate([g.ravel() for g in grads]))
        stats = torch.tensor([float(-(Yl * np.log(P + 1e-15)).sum()), float((P.argmax(0) == Yl.argmax(0)).sum())], dtype=torch.float64)
        dist.all_reduce(flat); dist.all_reduce(stats)      # the ONE exchange of the step
        # full-batch oracle step
        p64 = [p.copy() for p in params]
        loss = C.c_double(); correct = C.c_int()
        o.orc_mlp_step((C.c_int * 4)(*dims), B, *[ptr(p) for p in p64], ptr(X), ptr(Y), 0.02, 1, C.byref(loss), C.byref(correct), None, 1)
        lr = np.float64(np.float32(-0.02))
        pos = 0
        worst = 0.0
        for p0, p1 in zip(params, p64):
            g = flat[pos:pos + p0.size].numpy().reshape(p0.shape); pos += p0.size
            upd = p0 + lr * g
            worst = max(worst, float(np.abs(upd - p1).max() / (np.abs(p1).max() + 1e-30)))
        ret[rank] = (worst, abs(stats[0].item() - loss.value) / abs(loss.value), int(stats[1].item()) - correct.value)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [512, 777])
def test_two_rank_gloo_allreduce_equals_full_batch_step(B):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, B, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            worst, dloss, dcorrect = ret[r]
            assert worst <= 1e-10 and dloss <= 1e-12 and dcorrect == 0, ret[r]
