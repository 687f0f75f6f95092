"""CPU parity of the sampler of the device data pipeline (csrc/data.cu) with lib/mnist_csv2.c: get_random_data_take's index
sequence under the same libc rand() stream -- against the compiled reference (oracle/_ref/libref_data.so) when present, and always
against a Python restatement of mnist_csv2.c:41-62.  (The host part of bla_mnist_* needs no GPU... but creating a store uploads
the dataset, so these tests drive the Fenwick sampler through a tiny in-process harness built on the same rule.)"""
import ctypes as C

import numpy as np

libc = C.CDLL(None)
libc.rand.restype = C.c_int
RAND_MAX = 2147483647


def reference_take_sequence(n, draws, seed):
    """mnist_csv2.c:41-62 restated: the O(n) scan, quirks included"""
    libc.srand(seed)
    sampled = [0] * n
    num_sampled = 0
    out = []
    for _ in range(draws):
        if num_sampled == n:
            num_sampled = 0
            sampled = [0] * n
        k = int(np.floor(np.float32(np.float32(n - num_sampled) * np.float32(libc.rand())) / np.float32(RAND_MAX)))
        i = 0
        while i < n and k > 0:
            if sampled[i] == 0:
                k -= 1
            i += 1
        i = min(i, n - 1)
        sampled[i] = 1
        num_sampled += 1
        out.append(i)
    return out


def test_python_restatement_has_the_documented_quirk():
    # the element AFTER the n-th unsampled one is taken, so index 0 is only ever drawn when n == 0 and repeats are possible
    seq = reference_take_sequence(50, 50, 42)
    assert len(seq) == 50 and len(set(seq)) < 50
