"""Run under torchrun with >= 2 GPUs: 3 data-parallel SGD steps of the MLP (one process per GPU, NCCL all-reduce
inside bla_mlp_train_step) against the f64 oracle's full-batch steps.  Prints DP_CHECK_OK on rank 0."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import bla_b200 as b
    from helpers import load_oracle, ptr, rel_err
    dp = importlib.import_module("big-linear-algebra_b200.dp")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    b.bla_init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b.bla_comm_init(dp.exchange_unique_id(b, dist, rank, device="cuda"), rank, world)
    path = {"fp32": b.GEMM_FP32, "3xtf32": b.GEMM_3XTF32}[os.environ.get("DP_PATH", "fp32")]
    b.bla_set_gemm_path(path)
    B, steps = int(os.environ.get("DP_BATCH", "2000")), 3
    dims = (C.c_int * 4)(784, 256, 128, 10)
    rng = np.random.default_rng(5)                             # identical on every rank
    p32 = [np.ascontiguousarray(rng.uniform(-0.08, 0.08, s), np.float32) for s in ((256, 784), (256,), (128, 256), (128,), (10, 128), (10,))]
    off, cnt = dp.shard_columns(B, world, rank)
    net = b.bla_mlp_create(dims, cnt)
    b.bla_mlp_set_params(net, *[ptr(p) for p in p32])
    p64 = [p.astype(np.float64) for p in p32]
    o64 = load_oracle(np.float64)
    ok = True
    for s in range(steps):
        X = rng.integers(0, 256, (784, B)).astype(np.float32)
        labels = rng.integers(0, 10, B)
        Y = np.zeros((10, B), np.float32); Y[labels, np.arange(B)] = 1
        Xl, Yl = np.ascontiguousarray(X[:, off:off + cnt]), np.ascontiguousarray(Y[:, off:off + cnt])
        stats = np.zeros(2)
        b.bla_mlp_train_step(net, ptr(Xl), ptr(Yl), cnt, B, off, 0.02, ptr(stats))
        if rank == 0:
            loss = C.c_double(); correct = C.c_int()
            o64.orc_mlp_step(dims, B, *[ptr(p) for p in p64], ptr(X.astype(np.float64)), ptr(Y.astype(np.float64)), 0.02, 1,
                             C.byref(loss), C.byref(correct), None, 1)
            ok &= abs(stats[0] - loss.value) <= 1e-4 * abs(loss.value) and int(stats[1]) == correct.value
    # statistics accumulated over steps that do not read them: one collective read returns the global totals, counted once
    tot_loss = 0.0; tot_correct = 0
    for s in range(2):
        X = rng.integers(0, 256, (784, B)).astype(np.float32)
        labels = rng.integers(0, 10, B)
        Y = np.zeros((10, B), np.float32); Y[labels, np.arange(B)] = 1
        b.bla_mlp_train_step(net, ptr(np.ascontiguousarray(X[:, off:off + cnt])), ptr(np.ascontiguousarray(Y[:, off:off + cnt])), cnt, B, off, 0.02, None)
        if rank == 0:
            loss = C.c_double(); correct = C.c_int()
            o64.orc_mlp_step(dims, B, *[ptr(p) for p in p64], ptr(X.astype(np.float64)), ptr(Y.astype(np.float64)), 0.02, 1,
                             C.byref(loss), C.byref(correct), None, 1)
            tot_loss += loss.value; tot_correct += correct.value
    stats = np.zeros(2)
    b.bla_mlp_read_stats(net, ptr(stats))
    if rank == 0:
        ok &= abs(stats[0] - tot_loss) <= 1e-4 * abs(tot_loss) and int(stats[1]) == tot_correct
    steps += 2
    got = [np.empty_like(p) for p in p32]
    b.bla_mlp_get_params(net, *[ptr(g) for g in got])
    # every rank must hold identical parameters (same all-reduced gradient, same update)
    flat = torch.from_numpy(np.concatenate([g.ravel() for g in got])).cuda()
    ref = flat.clone(); dist.broadcast(ref, 0)
    same = bool(torch.equal(flat, ref))
    gathered = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([1.0 if same else 0.0], device="cuda"))
    if rank == 0:
        tol = 1e-5 if path == b.GEMM_FP32 else 1e-4
        errs = [rel_err(g, w) for g, w in zip(got, p64)]
        ok &= all(e <= tol * steps for e in errs) and all(t.item() == 1.0 for t in gathered)
        print("DP_CHECK_OK" if ok else "DP_CHECK_FAIL", "world", world, "errs", ["%.2e" % e for e in errs], flush=True)
    # device-resident batches: the step runs as a captured CUDA graph (csrc/mlp.cu step_graph) -- eager, captured and replayed steps
    # over two alternating resident batches against the oracle's same six full-batch steps
    net2 = b.bla_mlp_create(dims, cnt)
    b.bla_mlp_set_params(net2, *[ptr(p) for p in p32])
    q64 = [p.astype(np.float64) for p in p32]
    bufs = []
    for k in range(2):
        X = rng.integers(0, 256, (784, B)).astype(np.float32)
        labels = rng.integers(0, 10, B)
        Y = np.zeros((10, B), np.float32); Y[labels, np.arange(B)] = 1
        Xl, Yl = np.ascontiguousarray(X[:, off:off + cnt]), np.ascontiguousarray(Y[:, off:off + cnt])
        xd, yd = b.bla_malloc_device(Xl.nbytes), b.bla_malloc_device(Yl.nbytes)
        b.bla_copy_h2d(xd, ptr(Xl), Xl.nbytes); b.bla_copy_h2d(yd, ptr(Yl), Yl.nbytes); b.bla_sync()
        bufs.append((xd, yd, X, Y))
    launches0 = b.bla_launch_count()
    for k in range(6):
        xd, yd, X, Y = bufs[k % 2]
        b.bla_mlp_train_step(net2, xd, yd, cnt, B, off, 0.02, None)
        if rank == 0:
            loss = C.c_double(); correct = C.c_int()
            o64.orc_mlp_step(dims, B, *[ptr(p) for p in q64], ptr(X.astype(np.float64)), ptr(Y.astype(np.float64)), 0.02, 1,
                             C.byref(loss), C.byref(correct), None, 1)
    got2 = [np.empty_like(p) for p in p32]
    b.bla_mlp_get_params(net2, *[ptr(g) for g in got2])
    flat2 = torch.from_numpy(np.concatenate([g.ravel() for g in got2])).cuda()
    ref2 = flat2.clone(); dist.broadcast(ref2, 0)
    same2 = torch.tensor([1.0 if torch.equal(flat2, ref2) else 0.0], device="cuda")
    dist.all_reduce(same2, op=dist.ReduceOp.MIN)
    if rank == 0:
        errs2 = [rel_err(g, w) for g, w in zip(got2, q64)]
        tol = 1e-5 if path == b.GEMM_FP32 else 1e-4
        g_ok = all(e <= tol * 6 for e in errs2) and same2.item() == 1.0
        print("DP_GRAPH_OK" if g_ok else "DP_GRAPH_FAIL", "peer_windows", int(b.bla_comm_peer_windows()), "launches", int(b.bla_launch_count() - launches0),
              "errs", ["%.2e" % e for e in errs2], flush=True)
    for xd, yd, _, _ in bufs:
        b.bla_free(xd); b.bla_free(yd)
    b.bla_mlp_destroy(net2)
    # BASELINE.json configs[1] data parallel: every rank walks its own shard of the samples, the 10 x 784 gradient
    # and the sample count are all-reduced inside bla_hinge_iteration; three iterations against the float oracle on ALL samples
    o32 = load_oracle(np.float32)
    nh, F = 2500 * world + 37, 784
    hr = np.random.default_rng(11)
    hx = hr.integers(0, 256, (nh, F)).astype(np.float32)
    hl = hr.integers(0, 10, nh).astype(np.int32)
    hw0 = (hr.random((10, F)) / 10 - 0.05).astype(np.float32)
    ho, hc = dp.shard_columns(nh, world, rank)
    store = b.bla_mnist_from_arrays(ptr(np.ascontiguousarray(hx[ho:ho + hc])), ptr(np.ascontiguousarray(hl[ho:ho + hc].astype(np.float32))), hc, F)
    hg = b.bla_hinge_create(F, 10, hc)
    b.bla_hinge_set_weights(hg, ptr(hw0))
    hw_ref, hgrad_ref = hw0.copy(), np.zeros((10, F), np.float32)
    h_ok = True
    for it in range(3):
        norms = np.zeros(10, np.float32); norms_ref = np.zeros(10, np.float32)
        b.bla_hinge_iteration(hg, store, 0.001, ptr(norms))
        if rank == 0:
            o32.orc_hinge_iter(nh, F, ptr(hw_ref), ptr(hgrad_ref), ptr(hx), ptr(hl), C.c_float(0.001), ptr(norms_ref))
            h_ok &= bool(np.allclose(norms, norms_ref, rtol=1e-4, atol=1e-6))
    hgot = np.empty_like(hw0)
    b.bla_hinge_get_weights(hg, ptr(hgot))
    ht = torch.from_numpy(hgot.ravel().copy()).cuda()
    href = ht.clone(); dist.broadcast(href, 0)
    hsame = torch.tensor([1.0 if torch.equal(ht, href) else 0.0], device="cuda")
    dist.all_reduce(hsame, op=dist.ReduceOp.MIN)
    if rank == 0:
        h_ok &= rel_err(hgot, hw_ref) <= 1e-4 and hsame.item() == 1.0
        print("HINGE_DP_OK" if h_ok else "HINGE_DP_FAIL", "weights err %.2e" % rel_err(hgot, hw_ref), "identical on all ranks", bool(hsame.item()), flush=True)
    b.bla_hinge_destroy(hg); b.bla_mnist_destroy(store)
    # row-sharded GEMM (BASELINE.json configs[3]): every rank multiplies its row block by the broadcast B
    n = 512
    rowsn = n // world
    Ah = np.ascontiguousarray(np.random.default_rng(100 + rank).uniform(-0.5, 0.5, (rowsn, n)), np.float32)
    Bh = np.ascontiguousarray(np.random.default_rng(7).uniform(-0.5, 0.5, (n, n)), np.float32)      # only rank 0's copy is used
    Ad = b.bla_malloc_device(Ah.nbytes); Bd = b.bla_malloc_device(Bh.nbytes); Cd = b.bla_malloc_device(Ah.nbytes)
    b.bla_copy_h2d(Ad, ptr(Ah), Ah.nbytes)
    if rank == 0:
        b.bla_copy_h2d(Bd, ptr(Bh), Bh.nbytes)
    else:
        b.bla_memset_zero(Bd, Bh.nbytes)
    b.bla_broadcast_f32(Bd, n * n, 0)
    b.bla_gemm(0, 0, rowsn, n, n, Ad, n, Bd, n, Cd, n)
    Ch = np.empty_like(Ah); b.bla_copy_d2h(ptr(Ch), Cd, Ch.nbytes); b.bla_sync()
    gemm_ok = rel_err(Ch, Ah.astype(np.float64) @ Bh.astype(np.float64)) <= (1e-5 if path == b.GEMM_FP32 else 1e-4)
    flags = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if gemm_ok else 0.0], device="cuda"))
    if rank == 0:
        print("SHARDED_GEMM_OK" if all(f.item() == 1.0 for f in flags) else "SHARDED_GEMM_FAIL", flush=True)
    # U-Net (BASELINE.json configs[4]) data-parallel over images: every rank steps its shard, gradients are all-reduced inside
    # bla_unet_train_step; the result must equal the float64 full-batch gradient (tests/unet_ref.py) on every rank
    import unet_ref
    import test_unet_gpu as T
    b.bla_set_quirks(0)
    cfg, imgs_total = T.SMALL, 2 * world
    unet, tensors = T.make_net(b, cfg, imgs_total)
    flat = unet_ref.synthetic_params(cfg, tensors, imgs_total, 5)
    b.bla_unet_set_params(unet, ptr(flat))
    x, temb, noise = T.inputs(cfg, imgs_total, 3)                       # identical on every rank
    lo, hi = 2 * rank, 2 * rank + 2
    loss = np.zeros(1)
    b.bla_unet_train_step(unet, ptr(np.ascontiguousarray(x[lo:hi])), ptr(np.ascontiguousarray(temb[lo:hi])),
                          ptr(np.ascontiguousarray(noise[lo:hi])), 2, 0.0, ptr(loss))
    g = np.empty(flat.size, np.float32)
    b.bla_unet_get_grads(unet, ptr(g))
    _, want_loss, want_g = unet_ref.reference_step(cfg, tensors, flat, x, temb, noise, 0)
    tolu = 1e-5 if path == b.GEMM_FP32 else 1e-4
    worst = max(rel_err(g[o:o + c], want_g[o:o + c]) for _, o, c in tensors)
    unet_ok = worst <= tolu and abs(loss[0] - want_loss) <= 1e-4 * abs(want_loss)
    flags = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1.0 if unet_ok else 0.0], device="cuda"))
    if rank == 0:
        print("UNET_DP_OK" if all(f.item() == 1.0 for f in flags) else "UNET_DP_FAIL", "worst grad err %.2e" % worst, "loss", loss[0], want_loss,
              flush=True)
    b.bla_unet_destroy(unet)
    b.bla_set_quirks(1)
    b.bla_mlp_destroy(net)
    b.bla_comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
