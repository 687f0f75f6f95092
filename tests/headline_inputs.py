"""Seeded inputs shared by tests/golden/make_golden_headline.py (which feeds them to the compiled reference) and
tests/test_headline_gpu.py (which feeds them to libbla.so): the 60,000-column MNIST-shaped batch and the square GEMM operands of
BASELINE.json configs[2] / [3]."""
import numpy as np

DIMS = (784, 256, 128, 10)


def mlp_batch(B, seed=2026):
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 256, (DIMS[0], B), dtype=np.uint8).astype(np.float32)       # integer pixels, as lib/mnist_csv2.c:28 delivers them
    labels = rng.integers(0, DIMS[3], B)
    Y = np.zeros((DIMS[3], B), np.float32); Y[labels, np.arange(B)] = 1
    return X, Y


def mlp_params(seed=2027):
    rng = np.random.default_rng(seed)                                                # He-uniform ranges of model/mnist_nn.c:97-121, biases 0
    shapes = (((256, 784), 0.0875), ((256, 1), 0.0), ((128, 256), 0.153), ((128, 1), 0.0), ((10, 128), 0.2165), ((10, 1), 0.0))
    return [rng.uniform(-r, r, s).astype(np.float32) for s, r in shapes]


def splitmix_uniform(n, seed, lo, hi):
    """Host twin of csrc/kernels.h uniform_at (the generator behind bla_fill_uniform): bit-identical float32 values."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    # fmaf(hi - lo, u, lo) in float32: the product of a float32 pair is exact in float64, one rounding on the way back
    return (np.float64(np.float32(hi) - np.float32(lo)) * u.astype(np.float64) + np.float64(np.float32(lo))).astype(np.float32)


def gemm_inputs(n):
    return splitmix_uniform(n * n, 1, -0.5, 0.5).reshape(n, n), splitmix_uniform(n * n, 2, -0.5, 0.5).reshape(n, n)


def sample_index(size, count, seed):
    return np.random.default_rng(seed).integers(0, size, count)
