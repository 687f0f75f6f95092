#!/usr/bin/env python
"""bench.py -- headline benchmark of the big-linear-algebra B200 hot path.

Workload (BASELINE.json configs[2], the data-parallel one): one mini-batch SGD step of the
reference's MNIST MLP (model/mnist_nn.c:193-337; 784-256-128-10, ReLU/ReLU/softmax, CE loss) on a
GLOBAL batch of 60,000 synthetic MNIST-shaped samples, sharded column-wise over the N GPUs of one
box (one process per GPU, one NCCL all-reduce of the flat gradient buffer per step).  A "step" is
one pass of that hot loop.  `value` = samples/s with inputs resident in HBM; `e2e` = the same step
called through the C-ABI with pinned HOST buffers (H2D of the batch and D2H of the loss inside the
timed region).  `roofline` describes the dominant kernel (the layer-1 GEMM, 40 % of the step's
flops), timed alone with CUDA events on the library stream.  `extras` carries the other numbers
BASELINE.json's metric names: the square GEMM sweep (TFLOP/s, FP32 and 3xTF32 paths) and the
elementwise / norm kernels (GB/s against measured HBM bandwidth).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--impl reference times the reference's own single-threaded C (oracle/_ref, built from
/root/reference by oracle/build_ref.sh; falls back to the pinned restatement in oracle/) on the
host CPU on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS = (784, 256, 128, 10)
GLOBAL_BATCH = 60000
FLOP_PER_SAMPLE = 1007104          # fwd 469,504 + wgrad 469,504 + dgrad 68,096
LR = 0.02
METRIC = "mnist_mlp_train_samples_per_s"
UNIT = "samples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own C on the host
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """One MLP SGD step in the call order of model/mnist_nn.c:218-315, every matrix operation executed
    by the REFERENCE's compiled lib/matrix.c + lib/util.c (oracle/_ref/libref_f64.so); only relu' and the
    loss bookkeeping (model-local loops, O(B) work) are numpy.  Falls back to the pinned restatement
    (oracle/libbla_oracle_f64.so: orc_mlp_step) when oracle/_ref was not built."""

    def __init__(self):
        from helpers import load_oracle, load_ref, ref_available
        self.kind = "reference" if ref_available("f64") else "port"
        self.ref = load_ref("f64") if self.kind == "reference" else None
        self.orc = load_oracle(np.float64)
        rng = np.random.default_rng(42)
        self.params = [rng.uniform(-r, r, s) for s, r in (((256, 784), 0.0875), ((256, 1), 0.0), ((128, 256), 0.153),
                                                          ((128, 1), 0.0), ((10, 128), 0.2165), ((10, 1), 0.0))]

    def step(self, X, Y):
        if self.kind == "port":
            dims = (C.c_int * 4)(*DIMS)
            loss = C.c_double(); correct = C.c_int()
            self.orc.orc_mlp_step(dims, X.shape[1], *[p.ctypes.data_as(C.c_void_p) for p in self.params],
                                  X.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p), LR, 1, C.byref(loss),
                                  C.byref(correct), None, 1)
            return loss.value
        from helpers import as_matrix, matrix_to_numpy
        r = self.ref
        P = C.POINTER(r.MatrixT)
        W1, b1, W2, b2, W3, b3 = self.params
        Bn = X.shape[1]

        def mm(a, b):
            return r.matrix_multiply(a if isinstance(a, r.MatrixT) else a.contents, b if isinstance(b, r.MatrixT) else b.contents)

        def view(p):
            m = p.contents
            return np.ctypeslib.as_array(m.data, shape=(m.rows * m.cols,)).reshape(m.rows, m.cols)

        Xs = X.copy(); xm = as_matrix(r, Xs)
        r.matrix_scale(C.byref(xm), C.c_double(1 / np.float32(255.0)))
        acts, raws = [], []
        prev = xm
        for W, b, last in ((W1, b1, False), (W2, b2, False), (W3, b3, True)):
            z = mm(as_matrix(r, W), prev)
            bm = as_matrix(r, b)
            r.matrix_add_tile_columns(z, C.byref(bm))
            a = r.clone_matrix(z.contents)
            if last:
                r.softmax(a.contents.data, DIMS[3], Bn)
            else:
                r.relu(a.contents.data, a.contents.rows * Bn)
            raws.append(z); acts.append(a); prev = a
        A3 = view(acts[2])
        loss = float(-(Y.ravel() * np.log(A3.ravel() + 1e-15)).sum())
        self.last_correct = int((Y[np.argmax(A3, axis=0), np.arange(Bn)] == 1).sum())   # mnist_nn.c:237-247 (first maximum wins)
        g = r.clone_matrix(acts[2].contents)
        ym = as_matrix(r, Y)
        r.matrix_scale(C.byref(ym), C.c_double(-1.0)); r.matrix_add(g, C.byref(ym)); r.matrix_scale(C.byref(ym), C.c_double(-1.0))
        r.matrix_scale(g, C.c_double(1 / 784.0))
        grads = []
        below = [xm, acts[0], acts[1]]
        Ws = [W1, W2, W3]
        for l in (2, 1, 0):
            a_prev = below[l]
            ap = C.byref(a_prev) if isinstance(a_prev, r.MatrixT) else a_prev
            r.matrix_transpose(ap)
            dW = mm(g, a_prev)
            r.matrix_transpose(ap)
            db = r.matrix_col_sum(g.contents)
            grads.append((l, dW, db))
            if l > 0:
                wm = as_matrix(r, Ws[l])
                r.matrix_transpose(C.byref(wm))
                da = mm(wm, g)
                r.matrix_transpose(C.byref(wm))
                gz = r.clone_matrix(raws[l - 1].contents)
                v = view(gz); v[...] = (v > 0)                      # relu_ddx, model-local (mnist_nn.c:47-51)
                r.matrix_multiply_elementwise(gz, da)
                r.free_matrix(da); r.free_matrix(g)
                g = gz
        r.free_matrix(g)
        for l, dW, db in grads:
            for grad, param in ((dW, self.params[2 * l]), (db, self.params[2 * l + 1])):
                r.frobenius_norm(grad.contents)                     # clip_gradient (no-op threshold, :76-81)
                r.matrix_scale(grad, C.c_double(np.float32(-LR)))
                pm = as_matrix(r, param)
                r.matrix_add(C.byref(pm), grad)
                r.free_matrix(grad)
        for m in acts + raws:
            r.free_matrix(m)
        _ = (P, matrix_to_numpy)
        return loss


def synth_batch(B, seed):
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 256, (DIMS[0], B)).astype(np.float64)
    labels = rng.integers(0, DIMS[3], B)
    Y = np.zeros((DIMS[3], B)); Y[labels, np.arange(B)] = 1
    return X, Y


def time_cpu_reference(steps, warmup, budget_s):
    """Bounded sample: pick a batch so that (warmup + steps) CPU steps fit the budget."""
    cpu = CpuReference()
    X, Y = synth_batch(128, 1)
    t0 = time.perf_counter(); cpu.step(X, Y); probe = (time.perf_counter() - t0) / 128
    total = max(1, warmup + steps)
    B = int(budget_s / total / max(probe, 1e-9))
    B = max(64, min(8192, B // 64 * 64))
    X, Y = synth_batch(B, 2)
    for _ in range(warmup):
        cpu.step(X, Y)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu.step(X, Y)
    dt = time.perf_counter() - t0
    return dict(value=B * steps / dt, unit=UNIT, cores=1, kind=cpu.kind, ms_per_step=1e3 * dt / steps, columns=B,
                sample=f"{steps} SGD steps of {B} synthetic samples (of the 60,000-sample step), reference C single-threaded "
                       f"(it has no threads), gcc -O2, 1 of {os.cpu_count()} host cores")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warmup = min(args.warmup, 1)
    r = time_cpu_reference(steps, warmup, budget_s=90.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(world, args.scaling), sample_columns=r["columns"],
                           sample_note="the CPU arm times a bounded %d-column sample of the 60,000-column step (per-sample cost is flat in the batch width)" % r["columns"]),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(world, scaling="strong", allreduce=None):
    gb = GLOBAL_BATCH * world if scaling == "weak" else GLOBAL_BATCH
    return {"workload": "model/mnist_nn.c train step: MLP 784-256-128-10, one 60000-column batch per GPU "
                        "(data-parallel column shards, one NCCL all-reduce of the flat gradient per step)"
                        if scaling == "weak" else
                        "model/mnist_nn.c train step: MLP 784-256-128-10, global batch 60000 columns split over the GPUs",
            "global_batch": gb, "per_gpu_batch": gb // world, "parallelism": f"dp{world}",
            "allreduce": None if world == 1 else allreduce,
            "flop_per_sample": FLOP_PER_SAMPLE, "lr": LR}


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def summary(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        rows = [l.split(", ") for t, l in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [l.split(", ") for _, l in self.samples[-3:]]
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default=os.environ.get("BLA_BENCH_PATH", "auto"), choices=["auto", "fp32", "3xtf32"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default, BASELINE.json configs[2]): the global batch of 60,000 columns is split over the GPUs; "
                         "weak: every GPU takes a 60,000-column shard (global batch 60,000 x N)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    args.warmup = max(args.warmup, 3)

    import importlib
    import torch
    import bla_b200 as b
    dp = importlib.import_module("big-linear-algebra_b200.dp")
    if b.bla_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the big-linear-algebra B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    b.bla_init(local_rank)
    stream = torch.cuda.Stream()
    b.bla_set_stream(C.c_void_p(stream.cuda_stream))
    PATH = {"auto": b.GEMM_AUTO, "fp32": b.GEMM_FP32, "3xtf32": b.GEMM_3XTF32}[args.path]
    b.bla_set_gemm_path(PATH)

    dims = (C.c_int * 4)(*DIMS)
    shapes = ((DIMS[1], DIMS[0]), (DIMS[1],), (DIMS[2], DIMS[1]), (DIMS[2],), (DIMS[3], DIMS[2]), (DIMS[3],))

    def parity_batch(step, c0_, cnt):
        """Columns [c0_, c0_ + cnt) of the step's global synthetic batch (same generator on every rank)."""
        g = np.random.default_rng(9000 + step)
        X = g.integers(0, 256, (DIMS[0], GLOBAL_BATCH), dtype=np.uint8)[:, c0_:c0_ + cnt].astype(np.float32)
        labels = g.integers(0, DIMS[3], GLOBAL_BATCH)[c0_:c0_ + cnt]
        Y = np.zeros((DIMS[3], cnt), np.float32); Y[labels, np.arange(cnt)] = 1
        return np.ascontiguousarray(X), Y

    def parity_steps(c0_, cnt, nsteps=3):
        """nsteps SGD steps from He-uniform seed 42 on columns [c0_, c0_ + cnt) of the 60,000-column parity batches -> (params, loss sum)."""
        net_ = b.bla_mlp_create(dims, cnt)
        b.bla_mlp_init_params(net_, 42)
        init = [np.empty(sh, np.float32) for sh in shapes]
        b.bla_mlp_get_params(net_, *[g_.ctypes.data_as(C.c_void_p) for g_ in init])
        st = np.zeros(2); loss_sum = 0.0
        for k in range(nsteps):
            X, Y = parity_batch(k, c0_, cnt)
            b.bla_mlp_train_step(net_, X.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p), cnt, GLOBAL_BATCH, c0_, LR, st.ctypes.data_as(C.c_void_p))
            loss_sum += st[0]
        got = [np.empty(sh, np.float32) for sh in shapes]
        b.bla_mlp_get_params(net_, *[g_.ctypes.data_as(C.c_void_p) for g_ in got])
        b.bla_mlp_destroy(net_)
        return np.concatenate([g_.ravel() for g_ in got]), loss_sum, np.concatenate([g_.ravel() for g_ in init])

    # Data-parallel parity evidence for the line below (N > 1): rank 0 first runs three FULL-BATCH steps alone (no communicator yet) ...
    dp_single = parity_steps(0, GLOBAL_BATCH) if (world > 1 and rank == 0) else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        b.bla_comm_init(dp.exchange_unique_id(b, dist, rank, device="cuda"), rank, world)
    dp_parity = None
    if world > 1:
        # ... then all ranks run the same three steps on their column shards with the gradient all-reduce
        c0p, cntp = dp.shard_columns(GLOBAL_BATCH, world, rank)
        flat, loss_dp, init_flat = parity_steps(c0p, cntp)
        t = torch.from_numpy(flat).cuda()
        ref_t = t.clone(); dist.broadcast(ref_t, 0)
        same = torch.tensor([1.0 if torch.equal(t, ref_t) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if rank == 0:
            ref_flat, loss_single, _ = dp_single
            upd_dp, upd_single = flat.astype(np.float64) - init_flat, ref_flat.astype(np.float64) - init_flat
            dp_parity = {"steps": 3, "params_bit_identical_across_ranks": bool(same.item() == 1.0),
                         "max_rel_err_vs_single": float(np.linalg.norm(flat.astype(np.float64) - ref_flat) / np.linalg.norm(ref_flat.astype(np.float64))),
                         "update_rel_err_vs_single": float(np.linalg.norm(upd_dp - upd_single) / np.linalg.norm(upd_single)),
                         "loss_rel_err": float(abs(loss_dp - loss_single) / abs(loss_single)),
                         "what": "3 SGD steps from the same init: %d column shards + all-reduce against rank 0 alone on the full 60,000-column batch" % world}

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    Bg = GLOBAL_BATCH * world if args.scaling == "weak" else GLOBAL_BATCH
    c0, Bl = dp.shard_columns(Bg, world, rank)
    net = b.bla_mlp_create(dims, Bl)
    b.bla_mlp_init_params(net, 42)

    # resident synthetic batches, rotated so the inputs of consecutive steps never sit in L2 together
    x_bytes = DIMS[0] * Bl * 4
    nbuf = max(2, -(-(256 << 20) // x_bytes))
    Xd, Yd = [], []
    rng = np.random.default_rng(1234 + rank)
    for i in range(nbuf):
        xd = b.bla_malloc_device(x_bytes)
        # MNIST-shaped pixels: integers 0..255 stored as float32 (what the CSV loader hands the reference, mnist_csv2.c:28)
        px = rng.integers(0, 256, DIMS[0] * Bl, dtype=np.uint8)
        pd = b.bla_malloc_device(px.nbytes)
        b.bla_copy_h2d(pd, px.ctypes.data_as(C.c_void_p), px.nbytes)
        b.bla_u8_to_float(xd, pd, px.size, 1.0)
        b.bla_sync()
        b.bla_free(pd)
        labels = rng.integers(0, DIMS[3], Bl)
        Y = np.zeros((DIMS[3], Bl), np.float32); Y[labels, np.arange(Bl)] = 1
        yd = b.bla_malloc_device(Y.nbytes)
        b.bla_copy_h2d(yd, Y.ctypes.data_as(C.c_void_p), Y.nbytes)
        b.bla_sync()
        Xd.append(xd); Yd.append(yd)

    def step_resident(i):
        b.bla_mlp_train_step(net, Xd[i % nbuf], Yd[i % nbuf], Bl, Bg, c0, LR, None)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = b.bla_launch_count(); h0 = b.bla_h2d_bytes(); d0 = b.bla_d2h_bytes()
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return dict(ms=ms, launches=b.bla_launch_count() - l0, h2d=b.bla_h2d_bytes() - h0, d2h=b.bla_d2h_bytes() - d0, t0=t0, t1=t1)

    clocks = Clocks(local_rank) if rank == 0 else None
    time.sleep(0.3)
    # device-resident steps replay one captured graph per batch buffer: every buffer is seen eagerly once and captured once in the warm-up
    args.warmup = max(args.warmup, 2 * nbuf + 2)
    res = timed(step_resident, args.steps, args.warmup)
    clk = clocks.summary(res["t0"], res["t1"]) if clocks else None
    ms_per_step = res["ms"] / args.steps
    value = Bg / (ms_per_step * 1e-3)
    stats = np.zeros(2)
    b.bla_mlp_read_stats(net, stats.ctypes.data_as(C.c_void_p))
    # the same step for >= 2 s, under the power cap the box settles into (MEASURED_PEAKS.json: 1335 MHz median under sustained
    # tensor load): the clock record beside the short run's
    sustained = None
    if world == 1 and not args.no_extras:
        n_sus = int(max(args.steps, 2200.0 / ms_per_step))
        clocks2 = Clocks(local_rank)
        time.sleep(0.2)
        rs = timed(step_resident, n_sus, 3)
        sustained = {"steps": n_sus, "seconds": rs["ms"] * 1e-3, "ms_per_step": rs["ms"] / n_sus, "value": Bg / (rs["ms"] / n_sus * 1e-3),
                     "clocks": clocks2.summary(rs["t0"], rs["t1"])}
        b.bla_mlp_read_stats(net, stats.ctypes.data_as(C.c_void_p))

    # ---- e2e: pinned host float32 batch -> H2D -> step -> D2H of {loss, correct}, every step ----
    hx = b.bla_malloc_pinned(x_bytes)
    hy = b.bla_malloc_pinned(DIMS[3] * Bl * 4)
    hx_np = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(DIMS[0] * Bl,))
    hy_np = np.ctypeslib.as_array(C.cast(hy, C.POINTER(C.c_float)), shape=(DIMS[3], Bl))
    hx_np[:] = rng.integers(0, 256, DIMS[0] * Bl).astype(np.float32)
    labels = rng.integers(0, DIMS[3], Bl)
    hy_np[:] = 0; hy_np[labels, np.arange(Bl)] = 1
    e2e_stats = np.zeros(2)

    def step_e2e(i):
        b.bla_mlp_train_step(net, hx, hy, Bl, Bg, c0, LR, e2e_stats.ctypes.data_as(C.c_void_p))

    e2e_steps = max(3, min(args.steps, 10))
    e2e = timed(step_e2e, e2e_steps, 3)
    e2e_ms = e2e["ms"] / e2e_steps
    e2e_value = Bg / (e2e_ms * 1e-3)
    b.bla_mlp_set_host_chunking(net, 0)       # the same call with the batch staged in one piece, for the record
    e2e_one = timed(step_e2e, e2e_steps, 2)["ms"] / e2e_steps
    b.bla_mlp_set_host_chunking(net, -1)
    # ... and with the host-side byte packing of whole-number pixels off: the float chunks of round 1 (190.6 MB per step)
    e2e_float = None
    if hasattr(b, "bla_mlp_set_host_packing"):
        b.bla_mlp_set_host_packing(net, 0)
        e2e_f = timed(step_e2e, e2e_steps, 2)
        e2e_float = {"ms_per_step": e2e_f["ms"] / e2e_steps, "h2d_bytes_per_step": e2e_f["h2d"] // e2e_steps}
        b.bla_mlp_set_host_packing(net, -1)

    # the same end-to-end step from BYTE pixels (MNIST's native storage; additive entry point): 4x less PCIe traffic
    hx8 = b.bla_malloc_pinned(DIMS[0] * Bl)
    np.ctypeslib.as_array(C.cast(hx8, C.POINTER(C.c_ubyte)), shape=(DIMS[0] * Bl,))[:] = rng.integers(0, 256, DIMS[0] * Bl, dtype=np.uint8)

    def step_e2e_u8(i):
        b.bla_mlp_train_step_u8(net, hx8, hy, Bl, Bg, c0, LR, e2e_stats.ctypes.data_as(C.c_void_p))

    e2e8 = timed(step_e2e_u8, e2e_steps, 3)
    e2e8_ms = e2e8["ms"] / e2e_steps
    b.bla_mlp_set_host_chunking(net, 0)
    e2e8_one = timed(step_e2e_u8, e2e_steps, 2)["ms"] / e2e_steps
    b.bla_mlp_set_host_chunking(net, -1)

    # ---- roofline of the dominant kernel: layer-1 forward GEMM (256 x 784 x Bl), alone ----
    pk = peaks()
    a1 = b.bla_malloc_device(DIMS[1] * Bl * 4)
    w1 = b.bla_malloc_device(DIMS[1] * DIMS[0] * 4)
    b.bla_fill_uniform(w1, DIMS[1] * DIMS[0], 7, -0.08, 0.08)

    # The library issues this product as TWO launches when its last wave would be under half full (235 tile pairs on 74 CTA pairs):
    # a main launch over the columns of the full waves and a short split-K tail.  The dominant kernel is the main launch.
    n_main = int(b.bla_tc_main_columns(DIMS[1], Bl, DIMS[0])) if hasattr(b, "bla_tc_main_columns") else Bl
    if n_main <= 0 or n_main > Bl:
        n_main = Bl

    def gemm1(i):
        b.bla_gemm(0, 0, DIMS[1], n_main, DIMS[0], w1, DIMS[0], Xd[i % nbuf], Bl, a1, Bl)

    def gemm_layer(i):
        b.bla_gemm(0, 0, DIMS[1], Bl, DIMS[0], w1, DIMS[0], Xd[i % nbuf], Bl, a1, Bl)

    gr = timed(gemm1, 20, 5)
    gemm_ms = gr["ms"] / 20
    gemm_flop = 2.0 * DIMS[1] * DIMS[0] * n_main
    achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12
    layer_r = timed(gemm_layer, 20, 5)
    layer_ms = layer_r["ms"] / 20
    # the same launch on operands with a non-zero low part (uniform floats instead of integer pixels): all three MMAs per product
    xg = b.bla_malloc_device(DIMS[0] * Bl * 4)
    b.bla_fill_uniform(xg, DIMS[0] * Bl, 11, -1.0, 1.0)
    generic_ms = timed(lambda i: b.bla_gemm(0, 0, DIMS[1], n_main, DIMS[0], w1, DIMS[0], xg, Bl, a1, Bl), 20, 5)["ms"] / 20
    b.bla_free(xg)
    used_tc = b.bla_get_gemm_path() != b.GEMM_FP32 and tc_available(b)
    generic = None
    if used_tc:
        # MNIST pixels are integers 0..255: exact in TF32, so the low part of every X tile is zero and the kernel skips the W.lo(X)
        # product (gemm_tc.cu: the splitter warps flag all-zero lo tiles) -- TWO tcgen05 MMAs per product run in the step, three on
        # general fp32 data.  The peak is the one of what is issued.
        peak = pk["bf16"] / 2.0 / 2.0
        note = (f"3xTF32 on integer pixels: the lo(X) tile is all zero, so TWO of the three tcgen05 TF32 MMAs per product are issued; "
                f"peak = {pk['source']} bf16 burst {pk['bf16']} TFLOP/s / 2 (tf32) / 2 (MMAs per product); general fp32 operands "
                f"(three MMAs, peak / 3) in `generic_fp32_operands`")
        bound = "tensor"
        g_ach = gemm_flop / (generic_ms * 1e-3) / 1e12
        generic = {"achieved": g_ach, "peak": pk["bf16"] / 6.0, "unit": "TFLOP/s", "frac": g_ach / (pk["bf16"] / 6.0), "ms_per_launch": generic_ms,
                   "operands": "W1 and a uniform(-1, 1) X of the same shape: three MMAs per product"}
    else:
        peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        note = "FP32 FMA on the SIMT pipe: peak = 148 SMs x 128 lanes x 2 x clocks.max.sm (not in MEASURED_PEAKS.json)"
        bound = "fp32-simt"
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "kernel": "layer-1 forward GEMM, main launch 256x784x%d of the %d columns (38%% of step flops; the rest of the layer is a "
                          "tail launch over the remaining columns)" % (n_main, Bl),
                "ms_per_launch": gemm_ms, "peak_note": note,
                "whole_layer": {"shape": "256x784x%d" % Bl, "ms": layer_ms, "tflops": 2.0 * DIMS[1] * DIMS[0] * Bl / (layer_ms * 1e-3) / 1e12,
                                "launches": layer_r["launches"] // 20}}
    if generic:
        roofline["generic_fp32_operands"] = generic
        roofline["frac_of_nominal"] = achieved / (2250.0 / 4.0)   # nominal dense bf16 2250 TFLOP/s / 2 (tf32) / 2 (MMAs per product)
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (taken at the N=1 shape, tensor path)
    cap = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_dominant.json")
    if used_tc and Bl == 60000 and os.path.isfile(cap) and json.load(open(cap)).get("columns", Bl) == n_main:
        with open(cap) as f:
            cj = json.load(f)
        roofline["traffic"] = cj["dram_bytes_read"] + cj["dram_bytes_write"]
        roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum, %s; algorithmic minimum %.1f MB (X read once, "
                                    "A1 written once, W1)" % (cj["source"], (DIMS[0] * n_main + DIMS[1] * n_main + DIMS[0] * DIMS[1]) * 4 / 1e6))

    allreduce = None
    weak = None
    comm_ab = None
    if world > 1:
        allreduce = ("library kernel over NVLink peer windows (push + flags + rank-ordered sum + SGD update in one launch)"
                     if b.bla_comm_peer_windows() else "nccl")
        # A/B of the exchange itself: the 0.94 MB gradient all-reduce alone, and the same strong-scaling step with NCCL
        nflat = 235520
        gbuf = b.bla_malloc_device(nflat * 4)
        b.bla_fill_uniform(gbuf, nflat, 5, -1e-6, 1e-6)
        comm_ab = {"gradient_floats": nflat}
        had_peer = bool(b.bla_comm_peer_windows())
        for kind in (("peer_windows", 1), ("nccl", 0)) if had_peer else (("nccl", 0),):
            b.bla_comm_set_peer_windows(kind[1])
            r_ = timed(lambda i: b.bla_allreduce_sum_f32(gbuf, nflat), 300, 20)
            comm_ab[kind[0] + "_allreduce_us"] = r_["ms"] / 300 * 1e3
        if had_peer:                                   # still on NCCL: the headline step again, for the record
            r_ = timed(step_resident, min(args.steps, 100), 2 * nbuf + 2)
            comm_ab["ms_per_step_with_nccl"] = r_["ms"] / min(args.steps, 100)
            b.bla_comm_set_peer_windows(1)
        b.bla_free(gbuf)
        if args.scaling == "strong":
            # the weak curve beside the headline: every GPU takes a whole 60,000-column batch (global batch 60,000 x N)
            b.bla_mlp_destroy(net)
            for p_ in Xd + Yd:
                b.bla_free(p_)
            wnet = b.bla_mlp_create(dims, GLOBAL_BATCH)
            b.bla_mlp_init_params(wnet, 42)
            wx, wy = [], []
            for i in range(2):
                px = rng.integers(0, 256, DIMS[0] * GLOBAL_BATCH, dtype=np.uint8)
                pd = b.bla_malloc_device(px.nbytes); xd = b.bla_malloc_device(px.nbytes * 4)
                b.bla_copy_h2d(pd, px.ctypes.data_as(C.c_void_p), px.nbytes); b.bla_u8_to_float(xd, pd, px.size, 1.0); b.bla_sync(); b.bla_free(pd)
                labels = rng.integers(0, DIMS[3], GLOBAL_BATCH)
                Y = np.zeros((DIMS[3], GLOBAL_BATCH), np.float32); Y[labels, np.arange(GLOBAL_BATCH)] = 1
                yd = b.bla_malloc_device(Y.nbytes); b.bla_copy_h2d(yd, Y.ctypes.data_as(C.c_void_p), Y.nbytes); b.bla_sync()
                wx.append(xd); wy.append(yd)
            wsteps = min(args.steps, 100)
            wr = timed(lambda i: b.bla_mlp_train_step(wnet, wx[i % 2], wy[i % 2], GLOBAL_BATCH, GLOBAL_BATCH * world, GLOBAL_BATCH * rank, LR, None), wsteps, 6)
            weak = {"scaling": "weak", "global_batch": GLOBAL_BATCH * world, "per_gpu_batch": GLOBAL_BATCH, "ms_per_step": wr["ms"] / wsteps,
                    "value": GLOBAL_BATCH * world / (wr["ms"] / wsteps * 1e-3), "unit": UNIT}
            b.bla_mlp_destroy(wnet)
            for p_ in wx + wy:
                b.bla_free(p_)

    extras = None
    if not args.no_extras and rank == 0 and world == 1:
        extras = run_extras(b, torch, stream, pk)
    if not args.no_extras and world > 1:
        sharded = run_sharded_gemm(b, torch, dist, stream, rank, world, barrier)
        unet_dp = run_unet_dp(b, torch, dist, stream, world, barrier)
        if rank == 0:
            extras = {"gemm_sweep_row_sharded": sharded, "unet_data_parallel": unet_dp}
    if rank == 0 and weak is not None:
        extras = dict(extras or {}, weak_scaling=weak)
    if rank == 0 and comm_ab is not None:
        extras = dict(extras or {}, allreduce_ab=comm_ab)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = time_cpu_reference(2, 1, args.cpu_budget)
            cpu = {"value": r["value"], "unit": UNIT, "cores": 1, "kind": r["kind"], "sample": r["sample"]}
        except Exception as exc:   # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": dict(workload_config(world, args.scaling, allreduce), gemm_path=args.path,
                               l2=f"inputs rotate over {nbuf} resident batches ({nbuf * x_bytes >> 20} MiB > 126 MB L2)"),
                "tflops": value * FLOP_PER_SAMPLE / 1e12,
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"] // e2e_steps,
                        "d2h_bytes_per_step": e2e["d2h"] // e2e_steps, "ms_per_step": e2e_ms,
                        "launches_per_step": e2e["launches"] // e2e_steps, "ms_per_step_one_piece": e2e_one,
                        "float_chunks": e2e_float,
                        "api": "bla_mlp_train_step(host float32 X[784xB], Y[10xB], &stats): the batch is staged in column chunks behind "
                               "the training of the chunk before; chunks whose values are whole numbers 0..255 (checked bit for bit on "
                               "the host, every step, inside the timed region) cross PCIe as bytes and are widened on the device -- "
                               "h2d_bytes_per_step counts what was copied; `float_chunks` = the same call with that packing off"},
                "e2e_u8": {"value": Bg / (e2e8_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e8["h2d"] // e2e_steps,
                           "d2h_bytes_per_step": e2e8["d2h"] // e2e_steps, "ms_per_step": e2e8_ms,
                           "ms_per_step_one_piece": e2e8_one,
                           "api": "bla_mlp_train_step_u8(host uint8 X[784xB], Y[10xB], &stats)"},
                "gpu_launches": int(res["launches"]), "clocks": clk, "dp_parity": dp_parity, "sustained": sustained,
                "loss_per_sample_last": float(stats[0] / max(1, Bg * args.steps)) if world == 1 else None,
                "extras": extras}
        print(json.dumps(line), flush=True)
    if dist:
        b.bla_comm_destroy()
        dist.destroy_process_group()


def tc_available(b):
    return bool(getattr(b, "bla_tc_available", lambda: 0)())


def run_sharded_gemm(b, torch, dist, stream, rank, world, barrier):
    """BASELINE.json configs[3] on N GPUs: C[r] = A[r] . B with A and C row-sharded and B broadcast from rank 0 over
    NVLink (bla_broadcast_f32) -- no reduction.  Aggregate TFLOP/s, with and without the broadcast in the timed region."""
    out = []
    for n in (4096, 8192, 16384):
        rows = n // world
        A = b.bla_malloc_device(rows * n * 4); B = b.bla_malloc_device(n * n * 4); Cm = b.bla_malloc_device(rows * n * 4)
        b.bla_fill_uniform(A, rows * n, 100 + rank, -0.5, 0.5)
        if rank == 0:
            b.bla_fill_uniform(B, n * n, 2, -0.5, 0.5)
        row = {"n": n, "rows_per_gpu": rows}
        for name, path in (("3xtf32", b.GEMM_3XTF32), ("fp32", b.GEMM_FP32)):
            b.bla_set_gemm_path(path)
            for with_bcast in (False, True):
                def once():
                    if with_bcast:
                        b.bla_broadcast_f32(B, n * n, 0)
                    b.bla_gemm(0, 0, rows, n, n, A, n, B, n, Cm, n)
                b.bla_broadcast_f32(B, n * n, 0)
                for _ in range(2):
                    once()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 5 if n <= 8192 else 2
                e0.record(stream)
                for _ in range(iters):
                    once()
                e1.record(stream)
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row[name + ("_with_broadcast" if with_bcast else "") + "_tflops"] = 2.0 * n ** 3 / (float(t.item()) * 1e-3) / 1e12
        out.append(row)
        for p_ in (A, B, Cm):
            b.bla_free(p_)
    b.bla_set_gemm_path(b.GEMM_AUTO)
    return out


def run_unet_dp(b, torch, dist, stream, world, barrier, imgs=64):
    """BASELINE.json configs[4] on N GPUs, weak scaling: every rank steps its own 64 synthetic images, the 23.9 M-float gradient is
    all-reduced over NVLink inside bla_unet_train_step (the communicator is active).  Aggregate images/s, max over ranks."""
    uc = b.UnetConfig(32, (C.c_int * 4)(128, 256, 256, 256), 512, 3, 32, 16, 0.1, imgs, 7)
    net = b.bla_unet_create(C.byref(uc))
    b.bla_unet_init_params(net, 42)
    n3 = imgs * 3 * 32 * 32
    x = b.bla_malloc_device(n3 * 4); nz = b.bla_malloc_device(n3 * 4); te = b.bla_malloc_device(imgs * 512 * 4)
    b.bla_fill_uniform(x, n3, 1, -1, 1); b.bla_fill_uniform(nz, n3, 2, -1, 1); b.bla_fill_uniform(te, imgs * 512, 3, -1, 1)
    for _ in range(3):
        b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, None)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record(stream)
    for _ in range(iters):
        b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, None)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    b.bla_unet_destroy(net)
    for p_ in (x, nz, te):
        b.bla_free(p_)
    return {"workload": "model/cifar_unet.c train step, %d images per GPU, gradient all-reduce of 23.9 M floats per step" % imgs,
            "images_per_gpu": imgs, "n_gpus": world, "ms_per_step": ms, "images_per_s": imgs * world / (ms * 1e-3), "scaling": "weak"}


def run_extras(b, torch, stream, pk):
    """GEMM sweep (TFLOP/s) and elementwise / norm kernels (GB/s): the other numbers BASELINE.json's metric names."""
    out = {"gemm_sweep": [], "elementwise": []}

    def t(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # The HBM-bound kernels first, before the GEMM sweep: the 16384^3 products drive the GPU into its power cap and the SM clock stays
    # low for a while afterwards; the group-norm kernels (one latency chain per slab) are sensitive to it (measured in round 2:
    # 5.7 TB/s forward in a fresh process, 4.3 TB/s right after the sweep).
    n = 1 << 26                                  # 256 MiB per operand, well past the 126 MB L2
    rows, cols = 8192, 8192
    X = b.bla_matrix_device(rows, cols); Y = b.bla_matrix_device(rows, cols)
    b.bla_fill_uniform(C.cast(X.contents.data, C.c_void_p), n, 3, -1, 1); b.bla_fill_uniform(C.cast(Y.contents.data, C.c_void_p), n, 4, -1, 1)
    bias = b.bla_matrix_device(rows, 1)
    b.bla_fill_uniform(C.cast(bias.contents.data, C.c_void_p), rows, 5, -1, 1)
    cases = [("matrix_scale", 8, lambda: b.matrix_scale(X, C.c_float(1.0000001))),
             ("matrix_add", 12, lambda: b.matrix_add(X, Y)),
             ("matrix_multiply_elementwise", 12, lambda: b.matrix_multiply_elementwise(X, Y)),
             ("relu", 8, lambda: b.relu(C.cast(X.contents.data, C.c_void_p), n)),
             ("matrix_add_tile_columns", 8, lambda: b.matrix_add_tile_columns(X, bias)),
             ("matrix_col_sum", 4, lambda: b.free_matrix(b.matrix_col_sum(X.contents))),
             ("matrix_transpose(+copy back)", 16, lambda: b.matrix_transpose(X))]
    for name, bpe, fn in cases:
        ms = t(fn, 10)
        gbs = bpe * n / (ms * 1e-3) / 1e9
        out["elementwise"].append({"op": name, "bytes_per_elem": bpe, "gbs": gbs, "frac_hbm": gbs / pk["hbm"]})
    # group norm, U-Net shape batched: 256 images x 128 ch x 32x32, groups of 32 channels (8 / 12 B per element)
    imgs, Cn, HW = 256, 128, 1024
    ne = imgs * Cn * HW
    gx = b.bla_malloc_device(ne * 4); gy = b.bla_malloc_device(ne * 4); gd = b.bla_malloc_device(ne * 4)
    gv = b.bla_malloc_device(imgs * 4 * 4); gm = b.bla_malloc_device(imgs * 4 * 4)
    b.bla_fill_uniform(gx, ne, 6, -1, 2); b.bla_fill_uniform(gd, ne, 7, -1, 1)
    ms = t(lambda: b.bla_group_norm(gx, gy, gv, gm, imgs, Cn, HW, 32), 10)
    out["elementwise"].append({"op": "group_norm fwd 256x128x32x32", "bytes_per_elem": 8, "gbs": 8 * ne / (ms * 1e-3) / 1e9,
                               "frac_hbm": 8 * ne / (ms * 1e-3) / 1e9 / pk["hbm"]})
    ms = t(lambda: b.bla_group_norm_ddx(gd, gy, gx, gm, gv, imgs, Cn, HW, 32), 10)
    out["elementwise"].append({"op": "group_norm bwd 256x128x32x32", "bytes_per_elem": 12, "gbs": 12 * ne / (ms * 1e-3) / 1e9,
                               "frac_hbm": 12 * ne / (ms * 1e-3) / 1e9 / pk["hbm"]})
    out["hbm_peak_gbs"] = pk["hbm"]
    for p_ in (gx, gy, gd, gv, gm):
        b.bla_free(p_)
    for m_ in (X, Y, bias):
        b.free_matrix(m_)
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    saved = b.bla_get_gemm_path()
    for n in (1024, 2048, 4096, 8192, 16384):
        A = b.bla_malloc_device(n * n * 4); B = b.bla_malloc_device(n * n * 4); Cc = b.bla_malloc_device(n * n * 4)
        b.bla_fill_uniform(A, n * n, 1, -0.5, 0.5); b.bla_fill_uniform(B, n * n, 2, -0.5, 0.5)
        row = {"n": n}
        for name, path in (("fp32", b.GEMM_FP32), ("3xtf32", b.GEMM_3XTF32)):
            if path == b.GEMM_3XTF32 and not tc_available(b):
                continue
            b.bla_set_gemm_path(path)
            iters = 20 if n <= 2048 else (5 if n <= 8192 else 2)
            ms = t(lambda: b.bla_gemm(0, 0, n, n, n, A, n, B, n, Cc, n), iters)
            tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
            row[name + "_tflops"] = tf
            row[name + "_frac"] = tf / (fp32_peak if name == "fp32" else pk["bf16"] / 6.0)
            if name == "3xtf32":   # the measured cuBLAS bf16 figure is ~73 % of the nominal 2250 TFLOP/s: fractions above 1 are against it
                row["3xtf32_frac_of_nominal"] = tf / (2250.0 / 6.0)
        out["gemm_sweep"].append(row)
        for p in (A, B, Cc):
            b.bla_free(p)
    b.bla_set_gemm_path(saved)
    out["gemm_peaks"] = {"fp32_simt_tflops": fp32_peak, "3xtf32_tflops": pk["bf16"] / 6.0, "source": pk["source"]}

    # implicit-GEMM conv2d at the U-Net's shapes, batch of 64 images, both GEMM paths: 2*M*N*K flop
    out["conv"] = []
    imgs = 64
    for (Cn, H, F, k, st) in ((128, 32, 128, 3, 1), (128, 32, 256, 3, 2), (256, 16, 256, 3, 1), (256, 8, 256, 3, 1), (256, 16, 256, 1, 1)):
        Ho = -(-H // st)
        nx, nw, ny = imgs * Cn * H * H, F * Cn * k * k, imgs * F * Ho * Ho
        dxp = b.bla_malloc_device(nx * 4); dwp = b.bla_malloc_device(nw * 4); dyp = b.bla_malloc_device(ny * 4)
        gxp = b.bla_malloc_device(nx * 4); gwp = b.bla_malloc_device(nw * 4)
        b.bla_fill_uniform(dxp, nx, 8, -1, 1); b.bla_fill_uniform(dwp, nw, 9, -0.05, 0.05); b.bla_fill_uniform(dyp, ny, 10, -1, 1)
        flop = 2.0 * F * (imgs * Ho * Ho) * (Cn * k * k)
        row = {"shape": f"{imgs}x{Cn}x{H}x{H} -> {F}, k{k} s{st}", "gflop": flop / 1e9}
        ops = (("fprop", lambda: b.bla_conv2d_forward(dxp, dwp, dyp, imgs, Cn, H, H, F, k, st)),
               ("wgrad", lambda: b.bla_conv2d_wgrad(dxp, dyp, gwp, imgs, Cn, H, H, F, k, st)),
               ("dgrad", lambda: b.bla_conv2d_dgrad(dyp, dwp, gxp, imgs, Cn, H, H, F, k, st)))
        for pname, path, peak in (("fp32", b.GEMM_FP32, fp32_peak), ("3xtf32", b.GEMM_3XTF32, pk["bf16"] / 6.0)):
            if path == b.GEMM_3XTF32 and not tc_available(b):
                continue
            b.bla_set_gemm_path(path)
            for name, fn in ops:
                tc0 = b.bla_tc_launch_count()
                ms = t(fn, 5)
                if path == b.GEMM_3XTF32 and b.bla_tc_launch_count() == tc0:
                    continue                      # this op/shape has no tensor path yet: the fp32 figure stands
                row[f"{name}_{pname}_tflops"] = flop / (ms * 1e-3) / 1e12
                row[f"{name}_{pname}_frac"] = row[f"{name}_{pname}_tflops"] / peak
        b.bla_set_gemm_path(saved)
        out["conv"].append(row)
        for p_ in (dxp, dwp, dyp, gxp, gwp):
            b.bla_free(p_)
    out["unet"] = run_unet(b, t, pk)
    big = run_unet(b, t, pk, imgs=256)          # SURVEY 8(d) config 5 names batches of 1, 64 and 256 images
    out["unet"]["batch_256"] = {k: big[k] for k in ("images", "train_ms_per_step", "train_images_per_s", "train_tflops", "forward_ms",
                                                    "forward_images_per_s")}
    out["csv_codec"] = run_csv_codec(b)
    out["data_pipeline"] = run_data_pipeline(b)
    out["program_e2e"] = run_program_e2e()
    return out


def run_program_e2e(rows=5120):
    """The drop-in cost (BASELINE.json configs[2] 'model/mnist_nn.c relinks unchanged'): wall clock of the UNCHANGED reference
    program model/mnist_nn.c (SGD_BATCH_SIZE 512; `init`, then `train 1` over a synthetic MNIST-shaped CSV) linked against
    libbla.so, beside the same program linked against the reference's own lib/*.c (double build).  Process start, CUDA context
    creation, CSV parsing and the program's host-side loops (mnist_nn.c:38-91 on managed memory) are all inside the number."""
    import shutil
    import tempfile
    bin_dir = os.path.join(ROOT, "oracle", "_ref", "bin")
    progs = {"libbla": os.path.join(bin_dir, "bla_mnist_nn_b512"), "reference_f64": os.path.join(bin_dir, "ref_mnist_nn_f64_b512")}
    if not all(os.path.exists(p) for p in progs.values()):
        return {"unavailable": "oracle/_ref/bin programs not built (oracle/build_ref.sh needs /root/reference)"}
    res = {"program": "model/mnist_nn.c (SGD_BATCH_SIZE 512), `train 1`", "csv_rows": rows, "sgd_steps": rows // 512}
    tmp = tempfile.mkdtemp(prefix="bla_prog_")
    try:
        for sub in ("mnist", "mnist_nn"):
            os.makedirs(os.path.join(tmp, "data", sub))
        rng = np.random.default_rng(5)
        labels = rng.integers(0, 10, rows)
        proto = rng.integers(0, 256, (10, 784))
        with open(os.path.join(tmp, "data", "mnist", "mnist_train.csv"), "w") as f:
            for i in range(rows):
                px = np.clip(proto[labels[i]] + rng.integers(-60, 60, 784), 0, 255)
                f.write(f"{labels[i]}," + "".join(f"{int(v)}," for v in px) + "\n")
        env = dict(os.environ, BLA_PATH="auto")
        for name, exe in progs.items():
            subprocess.run([exe, "init"], cwd=tmp, capture_output=True, timeout=300, env=env)
            t0 = time.perf_counter()
            p = subprocess.run([exe, "train", "1"], cwd=tmp, capture_output=True, text=True, timeout=900, env=env)
            dt = time.perf_counter() - t0
            last = [l for l in p.stdout.splitlines() if l.startswith("Epoch")]
            res[name] = {"seconds": dt, "rc": p.returncode, "epoch_line": last[-1] if last else None}
        if res["libbla"]["rc"] == 0 and res["reference_f64"]["rc"] == 0:
            res["speedup"] = res["reference_f64"]["seconds"] / res["libbla"]["seconds"]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return res


def run_data_pipeline(b):
    """one epoch of model/mnist_nn.c:181-342 over a device-resident synthetic MNIST (60,000 x 784), sampler +
    gather + SGD steps, at the shipped batch size (64), at 512 and at one 60,000-sample batch; and the sampler alone beside the
    reference's O(examples) scan per draw (lib/mnist_csv2.c:41-62, compiled reference, a reported CPU baseline)."""
    n = 60000
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (n, 784)).astype(np.float32)
    y = rng.integers(0, 10, n).astype(np.float32)
    store = b.bla_mnist_from_arrays(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), n, 784)
    res = {"examples": n}
    dims = (C.c_int * 4)(*DIMS)
    for bs in (64, 512, 60000):
        net = b.bla_mlp_create(dims, bs)
        b.bla_mlp_init_params(net, 42)
        stats = np.zeros(2)
        b.bla_mlp_train_epoch(net, store, bs, LR, stats.ctypes.data_as(C.c_void_p))     # warm-up epoch
        t0 = time.perf_counter()
        b.bla_mlp_train_epoch(net, store, bs, LR, stats.ctypes.data_as(C.c_void_p))
        dt = time.perf_counter() - t0
        res[f"epoch_batch_{bs}"] = {"seconds": dt, "samples_per_s": n / dt, "steps": -(-n // bs)}
        b.bla_mlp_destroy(net)
    # BASELINE.json configs[1]: one full-batch iteration of the 10 one-vs-rest hinge classifiers over the
    # resident store (the sample matrix is streamed once: scores and gradients come from the same shared-memory tile)
    hg = b.bla_hinge_create(784, 10, n)
    w0 = (rng.random((10, 784)) / 10 - 0.05).astype(np.float32)
    b.bla_hinge_set_weights(hg, w0.ctypes.data_as(C.c_void_p))
    norms = np.zeros(10, np.float32)
    for _ in range(3):
        b.bla_hinge_iteration(hg, store, 0.001, None)
    b.bla_sync()
    t0 = time.perf_counter()
    for _ in range(20):
        b.bla_hinge_iteration(hg, store, 0.001, None)
    b.bla_hinge_iteration(hg, store, 0.001, norms.ctypes.data_as(C.c_void_p))
    dt = (time.perf_counter() - t0) / 21
    res["hinge_iteration"] = {"ms": dt * 1e3, "samples_per_s": n / dt, "gb_per_s_algorithmic": n * 784 * 4 / dt / 1e9,
                              "frac_hbm": n * 784 * 4 / dt / 1e9 / peaks()["hbm"],
                              "note": "algorithmic bytes = one pass over the 60,000 x 784 float samples (SURVEY 8d), which is what the kernel reads; "
                                      "20 flop per sample byte: the pass sits on the FP32 FMA / HBM ridge"}
    b.bla_hinge_destroy(hg)
    idx = np.empty(n, np.int32)
    b.bla_mnist_reset(store)
    t0 = time.perf_counter(); b.bla_mnist_sample_take(store, n, idx.ctypes.data_as(C.c_void_p)); dt = time.perf_counter() - t0
    res["sampler_draws_per_s"] = n / dt
    ref_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libref_data.so")
    if os.path.exists(ref_so):
        class MnistCSV(C.Structure):
            _fields_ = [("file", C.c_void_p), ("X", C.c_void_p), ("y", C.c_void_p), ("num_examples", C.c_int), ("num_sampled", C.c_int),
                        ("sampled", C.c_void_p)]

        class MnistExample(C.Structure):
            _fields_ = [("X", C.c_void_p), ("y", C.c_float), ("num_examples", C.c_int)]
        ref = C.CDLL(ref_so)
        ref.get_random_data_take.restype = MnistExample
        flags = np.zeros(n + 8, np.uint8)
        csv = MnistCSV(None, x.ctypes.data, y.ctypes.data, n, 0, flags.ctypes.data)
        draws = 6000                                        # the first tenth of an epoch: the scan gets longer as the epoch goes on
        t0 = time.perf_counter()
        for _ in range(draws):
            ref.get_random_data_take(C.byref(csv))
        res["reference_sampler_draws_per_s"] = draws / (time.perf_counter() - t0)
        res["reference_note"] = "lib/mnist_csv2.c get_random_data_take compiled from the reference, first %d draws of an epoch" % draws
    b.bla_mnist_destroy(store)
    return res


def run_csv_codec(b):
    """the checkpoint text format (lib/csv.c).  One 256 x 784 float tensor per file is the MLP's largest; timed on
    8 of them (1.6 M values, ~15 MB of text) from DEVICE memory through the pinned staging, beside the reference's own codec
    (oracle/_ref, compiled reference) on the same values from host memory -- a reported CPU baseline."""
    import tempfile
    rows, cols = 2048, 784
    n = rows * cols
    d = b.bla_malloc_device(n * 4)
    b.bla_fill_uniform(d, n, 11, -0.1, 0.1)
    res = {"values": n}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "w.csv").encode()
        t0 = time.perf_counter(); b.bla_csv_save(path, d, cols, rows); tw = time.perf_counter() - t0
        size = os.path.getsize(path)
        t0 = time.perf_counter(); b.bla_csv_load(path, d, n); tr = time.perf_counter() - t0
        res.update({"file_mb": size / 1e6, "save_mb_per_s": size / tw / 1e6, "load_mb_per_s": size / tr / 1e6,
                    "host_threads": os.cpu_count()})
        ref_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "libref_f32.so")
        if os.path.exists(ref_so):
            ref = C.CDLL(ref_so)
            ref.read_csv_contents.restype = C.c_void_p; ref.read_csv_contents.argtypes = [C.c_char_p]
            ref.write_csv_contents.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
            h = np.empty(n // 8, np.float32)              # an eighth of the values: the reference codec runs at ~10-20 MB/s
            b.bla_copy_d2h(h.ctypes.data_as(C.c_void_p), d, h.nbytes); b.bla_sync()
            rp = os.path.join(tmp, "r.csv").encode()
            t0 = time.perf_counter(); ref.write_csv_contents(rp, h.ctypes.data_as(C.c_void_p), cols, rows // 8); rw = time.perf_counter() - t0
            rsize = os.path.getsize(rp)
            t0 = time.perf_counter(); ref.read_csv_contents(rp); rr = time.perf_counter() - t0
            res.update({"reference_save_mb_per_s": rsize / rw / 1e6, "reference_load_mb_per_s": rsize / rr / 1e6,
                        "reference_note": "lib/csv.c compiled from the reference (oracle/_ref), single-threaded, %d values" % h.size})
    b.bla_free(d)
    return res


def run_unet(b, t, pk, imgs=64):
    """BASELINE.json configs[4]: model/cifar_unet.c (32x32x3, widths 128/256/256/256, 22 ResNet blocks, 5 attention blocks) as one
    batched device-resident training step -- forward, MSE against the noise, full backward, SGD -- on a synthetic batch.
    Algorithmic work: 7.13 GFLOP per image forward, 3x that for a training step.  The reference does one image
    at a time on one core: 30.6 s per forward+backward (probed with gcc -O2)."""
    uc = b.UnetConfig(32, (C.c_int * 4)(128, 256, 256, 256), 512, 3, 32, 16, 0.1, imgs, 7)
    net = b.bla_unet_create(C.byref(uc))
    b.bla_unet_init_params(net, 42)
    n3 = imgs * 3 * 32 * 32
    x = b.bla_malloc_device(n3 * 4); nz = b.bla_malloc_device(n3 * 4); te = b.bla_malloc_device(imgs * 512 * 4); o = b.bla_malloc_device(n3 * 4)
    b.bla_fill_uniform(x, n3, 1, -1, 1); b.bla_fill_uniform(nz, n3, 2, -1, 1); b.bla_fill_uniform(te, imgs * 512, 3, -1, 1)
    l0 = b.bla_launch_count()
    loss = np.zeros(1)
    b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, loss.ctypes.data_as(C.c_void_p))
    launches = int(b.bla_launch_count() - l0)
    step_ms = t(lambda: b.bla_unet_train_step(net, x, te, nz, imgs, 1e-6, None), 5, warm=2)
    fwd_ms = t(lambda: b.bla_unet_forward(net, x, te, imgs, o), 5, warm=2)
    peak = pk["bf16"] / 6.0
    res = {"workload": "model/cifar_unet.c train step, batch of %d synthetic 32x32x3 images, GEMM path auto (convs: 3xTF32 tcgen05), "
                       "BLA_QUIRKS=%d, dropout 0.1" % (imgs, b.bla_get_quirks()),
           "images": imgs, "params": int(b.bla_unet_num_params(net)), "launches_per_step": launches,
           "train_ms_per_step": step_ms, "train_images_per_s": imgs / (step_ms * 1e-3),
           "train_tflops": 3 * 7.13e9 * imgs / (step_ms * 1e-3) / 1e12, "train_frac_3xtf32_peak": 3 * 7.13e9 * imgs / (step_ms * 1e-3) / 1e12 / peak,
           "forward_ms": fwd_ms, "forward_images_per_s": imgs / (fwd_ms * 1e-3), "forward_tflops": 7.13e9 * imgs / (fwd_ms * 1e-3) / 1e12,
           "loss_per_image": float(loss[0]) / imgs,
           "reference_cpu_images_per_s": 1.0 / 30.6, "reference_cpu_note": "one image forward+backward, single-threaded, SURVEY 3.2 (probed)"}
    b.bla_unet_destroy(net)
    for p_ in (x, nz, te, o):
        b.bla_free(p_)
    return res


if __name__ == "__main__":
    main()
