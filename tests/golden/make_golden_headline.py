#!/usr/bin/env python
"""Golden values at the HEADLINE sizes, from the REAL reference (oracle/_ref/libref_f64.so, built by oracle/build_ref.sh from
/root/reference).  Run in the build container only (about three minutes of single-threaded reference C):

    bash oracle/build_ref.sh && python tests/golden/make_golden_headline.py

* one SGD step of the MNIST MLP (model/mnist_nn.c:218-315) on a 60,000-column batch -- BASELINE.json configs[2], the size
  bench.py reports -- driven through the reference's own lib/matrix.c / lib/util.c calls (bench.py: CpuReference);
* matrix_multiply (lib/matrix.c:35-57) at 1024^2 and 2048^2 -- BASELINE.json configs[3], SURVEY 8(d) config 4.
Inputs come from seeded generators that tests/test_headline_gpu.py repeats (the 188 MB batch is not stored); outputs are stored as
norms, checksums and sampled entries in tests/golden/headline.npz."""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from headline_inputs import gemm_inputs, mlp_batch, mlp_params, sample_index  # noqa: E402


def main():
    import bench
    from helpers import as_matrix, load_ref, matrix_to_numpy
    out = {}
    # ---- MLP step at 60,000 columns ----
    cpu = bench.CpuReference()
    assert cpu.kind == "reference", "oracle/_ref is not built"
    cpu.params = [p.astype(np.float64) for p in mlp_params()]
    X, Y = mlp_batch(60000)
    t0 = time.time()
    loss = cpu.step(X.astype(np.float64), Y.astype(np.float64))
    print("reference MLP step at 60,000 columns: %.1f s, loss %.6f, correct %d" % (time.time() - t0, loss, cpu.last_correct), flush=True)
    out["mlp_loss"] = np.float64(loss); out["mlp_correct"] = np.int64(cpu.last_correct)
    for i, p in enumerate(cpu.params):
        flat = p.ravel()
        out[f"mlp_p{i}_norm"] = np.float64(np.linalg.norm(flat)); out[f"mlp_p{i}_sum"] = np.float64(flat.sum())
        out[f"mlp_p{i}_sample"] = flat[sample_index(flat.size, 4096, 100 + i)]
    # ---- square GEMMs ----
    ref = load_ref("f64")
    for n in (1024, 2048):
        A, B = gemm_inputs(n)
        t0 = time.time()
        A64, B64 = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(B, np.float64)     # kept alive: the C side borrows the buffers
        c = ref.matrix_multiply(as_matrix(ref, A64), as_matrix(ref, B64))
        Cm = matrix_to_numpy(c, np.float64).reshape(n, n)
        print("reference matrix_multiply %d^2: %.1f s" % (n, time.time() - t0), flush=True)
        assert np.allclose(Cm[:4], A64[:4] @ B64, rtol=1e-9)                                     # the generator's own sanity check
        v = np.random.default_rng(7).uniform(-1, 1, n)
        out[f"gemm{n}_norm"] = np.float64(np.linalg.norm(Cm)); out[f"gemm{n}_Cv"] = Cm @ v; out[f"gemm{n}_vC"] = v @ Cm
        out[f"gemm{n}_sample"] = Cm.ravel()[sample_index(n * n, 4096, 200 + n)]
    np.savez_compressed(os.path.join(HERE, "headline.npz"), **out)
    print("wrote tests/golden/headline.npz")


if __name__ == "__main__":
    main()
