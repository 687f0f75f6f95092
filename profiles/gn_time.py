"""Times the batched group norm kernels (GB/s of algorithmic bytes: 8 B/elem forward, 12 B/elem backward)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bla_b200 as b
b.bla_init(0)
for imgs, Cn, HW in ((256, 128, 1024), (64, 128, 1024), (64, 256, 256), (64, 256, 64), (64, 256, 16), (256, 256, 256)):
    ne = imgs * Cn * HW
    G = Cn // 32
    x = b.bla_malloc_device(ne * 4); y = b.bla_malloc_device(ne * 4); d = b.bla_malloc_device(ne * 4)
    v = b.bla_malloc_device(imgs * G * 4); m = b.bla_malloc_device(imgs * G * 4)
    b.bla_fill_uniform(x, ne, 6, -1, 2); b.bla_fill_uniform(d, ne, 7, -1, 1)
    res = []
    for name, bpe, fn in (("fwd", 8, lambda: b.bla_group_norm(x, y, v, m, imgs, Cn, HW, 32)),
                          ("bwd", 12, lambda: b.bla_group_norm_ddx(d, y, x, m, v, imgs, Cn, HW, 32))):
        for _ in range(3): fn()
        b.bla_sync(); t0 = time.perf_counter()
        for _ in range(20): fn()
        b.bla_sync(); us = (time.perf_counter() - t0) / 20 * 1e6
        res.append(f"{name} {us:7.1f} us {bpe * ne / us / 1e3:6.0f} GB/s")
    print((imgs, Cn, HW), " | ".join(res), flush=True)
    for p in (x, y, d, v, m): b.bla_free(p)
