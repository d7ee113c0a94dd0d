"""GPU suite (-m gpu): the CUDA product, through its C ABI, against the oracle
(the real reference build oracle/_ref, prebuilt and shipped with the snapshot)
and the committed golden answers.  Bit-exact everywhere: integers, float bit
patterns of every butteraugli stage, JPEG bytes and the verbose trace."""
import numpy as np
import pytest

import parity
from guetzli_b200 import synth

pytestmark = pytest.mark.gpu

SIZES = [(64, 96, 7), (70, 51, 3), (136, 200, 4)]


@pytest.mark.parametrize("h,w,seed", SIZES)
def test_integer_stages(cuda_lib, ref, h, w, seed):
    parity.check_integer_stages(cuda_lib, ref, synth.gradnoise(h, w, seed))


@pytest.mark.parametrize("h,w,seed", SIZES)
def test_butteraugli_stages(cuda_lib, ref, h, w, seed):
    parity.check_butteraugli_stages(cuda_lib, ref, synth.gradnoise(h, w, seed))


@pytest.mark.parametrize("h,w,seed", SIZES + [(40, 33, 2)])
def test_compare_and_block_kernels(cuda_lib, ref, h, w, seed):
    parity.check_compare_and_blocks(cuda_lib, ref, synth.noise(h, w, seed))


@pytest.mark.parametrize("name", sorted(parity.GOLDEN))
def test_process_matches_golden(cuda_lib, name):
    st = parity.check_golden(cuda_lib, name)
    if not name.startswith("tiny"):
        assert st.device["gpu_launches"] > 0


def test_process_matches_reference_256(cuda_lib, ref):
    st = parity.check_process_vs_ref(cuda_lib, ref, synth.gradnoise(256, 256, 21), 92)
    assert st.device["gpu_launches"] > 0


@pytest.mark.parametrize("lookahead,new_model", [(3, False), (1, True), (5, False)])
def test_process_other_zeroing_params(cuda_lib, ref, lookahead, new_model):
    """zeroing_greedy_lookahead and the legacy zeroing score (processor.cc:391)."""
    parity.check_process_vs_ref(cuda_lib, ref, synth.gradnoise(96, 120, 12), 90, lookahead=lookahead,
                                new_zeroing_model=new_model)


def test_product_equals_port_on_512(cuda_lib, port_lib):
    """A larger case where the reference takes too long for a test: the CUDA
    product against the CPU restatement (itself pinned to the reference)."""
    rgb = synth.gradnoise(384, 512, 33)
    ok, jpeg, trace, _ = parity.run_process(cuda_lib, rgb, 90)
    ok2, jpeg2, trace2, _ = parity.run_process(port_lib, rgb, 90)
    assert ok and ok2 and trace == trace2 and jpeg == jpeg2


def test_full_size_properties_1080p(cuda_lib):
    """BASELINE config sizes, size-independent properties: (1) Compare of the
    unquantised image is deterministic across two contexts, (2) the distance is
    monotone under coarser quantisation, (3) scatter(restore) returns the exact
    distmap, (4) quantise is idempotent."""
    import guetzli_b200 as gb
    rgb = synth.gradnoise(1080, 1920, 4321)
    a = gb.DeviceImage(rgb, lib=cuda_lib)
    d0 = a.compare()
    dm0 = a.distmap()
    b = gb.DeviceImage(rgb, lib=cuda_lib)
    assert b.compare() == d0 and parity.bits_equal(b.distmap(), dm0)
    b.close()
    q2 = np.full((3, 64), 2, dtype=np.int32)
    q8 = np.full((3, 64), 8, dtype=np.int32)
    a.apply_global_quant(q2)
    c2 = a.download_candidate()
    d2 = a.compare()
    a.apply_global_quant(q8)
    d8 = a.compare()
    assert d0 <= d2 <= d8
    a.apply_global_quant(q2)
    assert np.array_equal(a.download_candidate(), c2)
    # zero a few coefficients then restore them: distmap must come back bit for bit
    dm2 = (a.compare(), a.distmap())
    flat = c2.reshape(-1)
    nz = np.flatnonzero(flat)[::5000][:200].astype(np.int32)
    a.scatter(nz, np.zeros(len(nz), dtype=np.int16))
    assert a.compare() >= 0
    a.scatter(nz, flat[nz])
    assert a.compare() == dm2[0] and parity.bits_equal(a.distmap(), dm2[1])
    a.close()


def _large_cases():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_large.json")
    return json.load(open(path)) if os.path.exists(path) else {}


@pytest.mark.parametrize("name", sorted(_large_cases()))
def test_process_matches_golden_full_size(cuda_lib, name):
    """BASELINE.json full-size configurations against the reference's own answers
    (tests/golden/make_golden_large.py: minutes to tens of minutes of CPU each)."""
    import hashlib
    g = _large_cases()[name]
    gens = {
        "noise1080p_s1234_q95": lambda: synth.noise(1080, 1920, 1234),
        "gradnoise4k_s4321_q90": lambda: synth.gradnoise(2160, 3840, 4321),
        "gradnoise1024_s1000_q84": lambda: synth.gradnoise(1024, 1024, 1000),
    }
    rgb = gens[name]()
    assert synth.sha256(rgb) == g["input_sha256"]
    ok, jpeg, trace, st = parity.run_process(cuda_lib, rgb, g["quality"])
    assert ok and len(jpeg) == g["jpeg_size"]
    assert hashlib.sha256(jpeg).hexdigest() == g["jpeg_sha256"]
    assert hashlib.sha256(trace.encode()).hexdigest() == g["trace_sha256"]
    assert [st.counters["number of iterations"], st.counters["number of iterations up"],
            st.counters["number of iterations down"]] == g["iterations"]


@pytest.mark.parametrize("name,world", [("bees_444x258_q95", 3), ("odd_70x51_s3_q88", 2)])
def test_strip_mode_on_one_gpu_matches_golden(cuda_lib, name, world):
    """Row-strip mode with the CUDA strip kernels (row-range launches, per-block
    exchange) driven by host threads that share this GPU: same bytes as untiled."""
    import hashlib
    import guetzli_b200 as gb
    g = parity.GOLDEN[name]
    rgb = parity.golden_input(name)
    h, w, _ = rgb.shape
    p = gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(g["quality"], lib=cuda_lib))
    ok, jpeg = gb.process_tiled_threads(p, rgb, w, h, world, lib=cuda_lib)
    assert ok and hashlib.sha256(jpeg).hexdigest() == g["jpeg_sha256"]
