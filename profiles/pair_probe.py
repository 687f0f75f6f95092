"""What bounds the CTA-pair 3xTF32 kernel?  Times the MLP's two long GEMMs (layer-1 forward main launch, layer-1 weight gradient)
with integer-pixel B (exact in TF32: the lo.hi product is skipped) and random B, under the kernel's experiment switches.
Usage: python profiles/pair_probe.py            (spawns one process per BLA_TC_DEBUG value)"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [("fwd1_main NN 256x56832x784", 0, 0, 256, 56832, 784), ("wgrad1 NT 256x784x60000", 0, 1, 256, 784, 60000),
         ("fwd2 NN 128x60000x256", 0, 0, 128, 60000, 256), ("dgrad1 TN 256x60000x128", 1, 0, 256, 60000, 128),
         ("wgrad2 NT 128x256x60000", 0, 1, 128, 256, 60000), ("square NN 8192", 0, 0, 8192, 8192, 8192)]


def child():
    import numpy as np
    import torch
    import bla_b200 as b
    b.bla_init(0)
    stream = torch.cuda.Stream()
    b.bla_set_stream(C.c_void_p(stream.cuda_stream))
    b.bla_set_gemm_path(b.GEMM_3XTF32)
    out = {}
    for name, ta, tb, M, N, K in CASES:
        A = b.bla_malloc_device(M * K * 4); B = b.bla_malloc_device(K * N * 4); Cm = b.bla_malloc_device(M * N * 4)
        b.bla_fill_uniform(A, M * K, 1, -0.5, 0.5)
        for kind in ("int", "rand"):
            if kind == "int":
                px = np.random.default_rng(0).integers(0, 256, K * N, dtype=np.uint8)
                pd = b.bla_malloc_device(px.nbytes)
                b.bla_copy_h2d(pd, px.ctypes.data_as(C.c_void_p), px.nbytes)
                b.bla_u8_to_float(B, pd, px.size, 1.0)
                b.bla_sync(); b.bla_free(pd)
            else:
                b.bla_fill_uniform(B, K * N, 2, -0.5, 0.5)
            fn = lambda: b.bla_gemm(ta, tb, M, N, K, A, M if ta else K, B, K if tb else N, Cm, N)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            out[f"{name} B={kind}"] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
        for p_ in (A, B, Cm):
            b.bla_free(p_)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    table = {}
    for dbg in sys.argv[1:] or ["0", "4", "5", "1"]:
        env = dict(os.environ, BLA_TC_DEBUG=dbg)
        p = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
        table[dbg] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-500:]}
    keys = list(next(iter(table.values())).keys())
    print("us per launch; BLA_TC_DEBUG: 0 = product, 4 = never skip a zero lo tile, 5 = no split work + three products, 1 = no split work + hi.hi only")
    print(f"{'case':40s}" + "".join(f"{'dbg ' + d:>10s}" for d in table))
    for k in keys:
        print(f"{k:40s}" + "".join(f"{table[d].get(k, float('nan')):10.1f}" for d in table))
