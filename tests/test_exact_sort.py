"""The prefix-exact replay of libstdc++'s std::sort (guetzli_b200/csrc/exact_sort.h)
must reproduce std::sort's arrangement -- including the order of equal keys -- on
the requested prefix."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("n,levels,seed", [(10, 3, 0), (17, 2, 1), (1000, 5, 2), (5000, 50, 3),
                                            (70000, 200, 4), (300000, 1000, 5), (200000, 3, 6),
                                            (65536, 1, 7), (100000, 10 ** 6, 8)])
def test_partial_sort_matches_std_sort(port_lib, n, levels, seed):
    rng = np.random.default_rng(seed)
    keys = (rng.integers(0, levels, n) / 7.0).astype(np.float32)
    if seed % 2:
        keys = np.sort(keys)[::-1].copy() if seed % 4 == 1 else np.sort(keys).copy()
    blocks = np.arange(n, dtype=np.int32)
    port_lib.gb200_debug_partial_sort.restype = C.c_size_t
    port_lib.gb200_debug_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    port_lib.gb200_debug_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rb, rk = blocks.copy(), keys.copy()
    port_lib.gb200_debug_std_sort(rb.ctypes.data, rk.ctypes.data, n)
    for want in sorted({1, 5, 16, 17, n // 100 + 1, n // 3 + 1, n}):
        b, k = blocks.copy(), keys.copy()
        k_end = port_lib.gb200_debug_partial_sort(b.ctypes.data, k.ctypes.data, n, want)
        assert min(want, n) <= k_end <= n
        assert np.array_equal(b[:k_end], rb[:k_end]), (n, want, k_end)
        assert np.array_equal(k[:k_end], rk[:k_end])
