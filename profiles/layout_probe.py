"""TMA row-rate hypothesis: one shape (256 x 57344 x 768) in the four operand layouts.  Rows TMA moves per 16-deep k-block and CTA of
a pair: K-major A 128 x 64 B, MN-major A 64 x 128 B, K-major B half 128 x 64 B, MN-major B half 64 x 128 B.
Usage: python profiles/layout_probe.py  (one process per BLA_TC_DEBUG value)"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import bla_b200 as b
    b.bla_init(0)
    stream = torch.cuda.Stream()
    b.bla_set_stream(C.c_void_p(stream.cuda_stream))
    b.bla_set_gemm_path(b.GEMM_3XTF32)
    out = {}
    for (M, N, K) in ((256, 57344, 768), (4096, 4096, 4096)):
        A = b.bla_malloc_device(M * K * 4); B = b.bla_malloc_device(K * N * 4); Cm = b.bla_malloc_device(M * N * 4)
        b.bla_fill_uniform(A, M * K, 1, -0.5, 0.5); b.bla_fill_uniform(B, K * N, 2, -0.5, 0.5)
        for ta in (0, 1):
            for tb in (0, 1):
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
                out[f"{M}x{N}x{K} {'T' if ta else 'N'}{'T' if tb else 'N'}"] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
        for p_ in (A, B, Cm):
            b.bla_free(p_)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    table = {}
    for dbg in sys.argv[1:] or ["0", "1"]:
        p = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BLA_TC_DEBUG=dbg), capture_output=True, text=True, timeout=300)
        table[dbg] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-500:]}
    keys = list(next(iter(table.values())).keys())
    print("us per launch; BLA_TC_DEBUG 0 = product kernel, 1 = no split work + hi.hi only (feed-bound floor)")
    print(f"{'case':34s}" + "".join(f"{'dbg ' + d:>10s}" for d in table))
    for k in keys:
        print(f"{k:34s}" + "".join(f"{table[d].get(k, float('nan')):10.1f}" for d in table))
