"""Row-strip mode (one image tiled over ranks, comm.h) on the CPU port: strips of
block rows + 56-row halo, per-block results exchanged; output must equal the
untiled result, i.e. the reference's bytes."""
import hashlib

import pytest

import guetzli_b200 as gb
import parity


@pytest.mark.parametrize("name,world", [("gradnoise_128x128_s11_q84", 2), ("odd_70x51_s3_q88", 3),
                                        ("gray_64x64_s9_q90", 2), ("bees_444x258_q95", 3),
                                        # strips narrower than the 7-block-row halo: 8 ranks on 16 block rows
                                        ("gradnoise_128x128_s11_q84", 8), ("odd_70x51_s3_q88", 4)])
def test_strip_mode_matches_golden(port_lib, name, world):
    g = parity.GOLDEN[name]
    rgb = parity.golden_input(name)
    h, w, _ = rgb.shape
    p = gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(g["quality"], lib=port_lib))
    ok, jpeg = gb.process_tiled_threads(p, rgb, w, h, world, lib=port_lib)
    assert ok and len(jpeg) == g["jpeg_size"]
    assert hashlib.sha256(jpeg).hexdigest() == g["jpeg_sha256"]
