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


@pytest.mark.parametrize("n,levels,seed", [(40000, 50, 1), (200000, 7, 2), (300000, 10 ** 6, 3), (150000, 1, 4),
                                            (120000, 3, 5)])
def test_device_partition_replay_matches_host(port_lib, n, levels, seed):
    """order_exact.h (experimental): introsort's partition passes over large ranges as
    parallel kernels must leave the same prefix as the host replay."""
    import guetzli_b200 as gb
    from guetzli_b200 import synth
    rng = np.random.default_rng(seed)
    keys = (rng.integers(0, levels, n) / 7.0).astype(np.float32)
    if seed == 5:
        keys = np.sort(keys)
    blocks = np.arange(n, dtype=np.int32)
    img = gb.DeviceImage(synth.gradnoise(16, 16, 1), lib=port_lib)
    port_lib.gb200_debug_partial_sort.restype = C.c_size_t
    port_lib.gb200_debug_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    port_lib.gb200_debug_device_partial_sort.restype = C.c_size_t
    port_lib.gb200_debug_device_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    for want in (1, 100, n // 50, n // 3, n):
        b0, k0 = blocks.copy(), keys.copy()
        ke0 = port_lib.gb200_debug_partial_sort(b0.ctypes.data, k0.ctypes.data, n, want)
        b1, k1 = blocks.copy(), keys.copy()
        ke1 = port_lib.gb200_debug_device_partial_sort(img._h, b1.ctypes.data, k1.ctypes.data, n, want)
        assert ke1 == ke0, (n, want)
        assert np.array_equal(b1[:ke1], b0[:ke0]) and np.array_equal(k1[:ke1], k0[:ke0]), (n, want)
    img.close()


def test_device_order_end_to_end_in_check_mode(port_lib):
    """GB200_DEVICE_ORDER=check (experimental path, order_exact.h): every reference-ordered
    iteration builds its candidate list and replays the large partition passes with the
    device functors as well and compares with the host replay; the host range threshold is
    lowered so that small images reach those passes.  Separate process: the switches are
    read once."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, 'tests')\n"
        "import numpy as np, parity, guetzli_b200 as gb\n"
        "from guetzli_b200 import synth\n"
        "lib = gb.load_library('oracle/_build/libguetzli_port.so')\n"
        "parity.check_golden(lib, 'gradnoise_128x128_s11_q84')\n"
        "parity.check_golden(lib, 'noise_48x40_s5_q95')\n"
        "rgb = np.ascontiguousarray(np.tile(synth.noise(64, 64, 3), (1, 2, 1)))\n"
        "a = parity.run_process(lib, rgb, 93)\n"
        "print('OK', len(a[1]))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GB200_DEVICE_ORDER="check", GB200_ORDER_HOST_RANGE="512", GB200_TIE_DEBUG="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert b"OK" in r.stdout and b"checked against the host replay" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n,levels,seed", [(300000, 10 ** 6, 3), (200000, 7, 2), (150000, 1, 4), (2000000, 10 ** 7, 9)])
def test_device_partition_replay_matches_host_cuda(cuda_lib, n, levels, seed):
    """The same comparison with the partition passes as CUDA kernels (tools/check_device_order.py)."""
    import guetzli_b200 as gb
    from guetzli_b200 import synth
    rng = np.random.default_rng(seed)
    keys = (rng.integers(0, levels, n) / 7.0).astype(np.float32)
    blocks = np.arange(n, dtype=np.int32)
    img = gb.DeviceImage(synth.gradnoise(16, 16, 1), lib=cuda_lib)
    cuda_lib.gb200_debug_partial_sort.restype = C.c_size_t
    cuda_lib.gb200_debug_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    cuda_lib.gb200_debug_device_partial_sort.restype = C.c_size_t
    cuda_lib.gb200_debug_device_partial_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    for want in (100, n // 50, n // 3):
        b0, k0 = blocks.copy(), keys.copy()
        ke0 = cuda_lib.gb200_debug_partial_sort(b0.ctypes.data, k0.ctypes.data, n, want)
        b1, k1 = blocks.copy(), keys.copy()
        ke1 = cuda_lib.gb200_debug_device_partial_sort(img._h, b1.ctypes.data, k1.ctypes.data, n, want)
        assert ke1 == ke0, (n, want)
        assert np.array_equal(b1[:ke1], b0[:ke0]) and np.array_equal(k1[:ke1], k0[:ke0]), (n, want)
    img.close()
