#!/usr/bin/env python
"""GPU check of the experimental device order replay (order_exact.h): the partition passes as
CUDA kernels against the host replay on synthetic keys, then one image end to end with
GB200_DEVICE_ORDER=check (set by the caller together with GB200_ORDER_HOST_RANGE)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

import guetzli_b200 as gb  # noqa: E402
from guetzli_b200 import synth  # noqa: E402

lib = gb.load_library()
lib.gb200_debug_partial_sort.restype = C.c_size_t
lib.gb200_debug_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
lib.gb200_debug_device_partial_sort.restype = C.c_size_t
lib.gb200_debug_device_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
img = gb.DeviceImage(synth.gradnoise(16, 16, 1))
bad = 0
for n, levels, seed in [(300000, 10 ** 6, 3), (200000, 7, 2), (150000, 1, 4), (2000000, 10 ** 7, 9)]:
    rng = np.random.default_rng(seed)
    keys = (rng.integers(0, levels, n) / 7.0).astype(np.float32)
    blocks = np.arange(n, dtype=np.int32)
    for want in (100, n // 50, n // 3):
        b0, k0 = blocks.copy(), keys.copy()
        ke0 = lib.gb200_debug_partial_sort(b0.ctypes.data, k0.ctypes.data, n, want)
        b1, k1 = blocks.copy(), keys.copy()
        ke1 = lib.gb200_debug_device_partial_sort(img._h, b1.ctypes.data, k1.ctypes.data, n, want)
        ok = ke1 == ke0 and np.array_equal(b1[:ke1], b0[:ke0]) and np.array_equal(k1[:ke1], k0[:ke0])
        bad += not ok
        print(n, levels, want, ke0, ke1, "ok" if ok else "MISMATCH")
img.close()
rgb = np.ascontiguousarray(np.tile(synth.noise(64, 64, 3), (1, 2, 1)))
st = gb.ProcessStats()
ok, jpeg = gb.process(gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(93)), st, rgb, 128, 64)
print("process", ok, len(jpeg), st.device["order_exact"], st.device["order_partial"])
sys.exit(1 if bad or not ok else 0)
