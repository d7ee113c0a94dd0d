#!/usr/bin/env python3
"""Generates tests/golden/golden.json (+ bees_rgb.npz) by running the UNMODIFIED
reference (oracle/_ref/libguetzli_ref.so, built by oracle/Makefile from
/root/reference) on named inputs.  Only runs where /root/reference exists; the
outputs are committed so that every test box can check against them.

The reference's own golden file (tests/golden_checksums.txt) cannot pin this tree
offline, so these hashes -- the reference's own outputs in this
container -- are the known answers for the path.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import reflib  # noqa: E402
from guetzli_b200 import synth  # noqa: E402


def gray(h, w, seed):
    g = synth.gradnoise(h, w, seed)[..., 1]
    return np.stack([g, g, g], axis=-1)


CASES = [
    ("gradnoise_64x96_s7_q90", lambda: synth.gradnoise(64, 96, 7), 90),
    ("noise_48x40_s5_q95", lambda: synth.noise(48, 40, 5), 95),
    ("tiny_20x40_s5_q95", lambda: synth.gradnoise(20, 40, 5), 95),
    ("odd_70x51_s3_q88", lambda: synth.gradnoise(70, 51, 3), 88),
    ("gray_64x64_s9_q90", lambda: gray(64, 64, 9), 90),
    ("gradnoise_128x128_s11_q84", lambda: synth.gradnoise(128, 128, 11), 84),
    ("min_32x32_s2_q90", lambda: synth.noise(32, 32, 2), 90),
    ("small_33x47_s4_q95", lambda: synth.gradnoise(33, 47, 4), 95),
    ("wide_32x200_s6_q88", lambda: synth.gradnoise(32, 200, 6), 88),
    ("flat_40x40_q95", lambda: np.full((40, 40, 3), 77, dtype=np.uint8), 95),
    ("bees_444x258_q95", None, 95),
]


def main():
    out = {}
    for name, gen, q in CASES:
        if gen is None:
            from PIL import Image
            rgb = np.array(Image.open("/root/reference/tests/bees.png").convert("RGB"))
            np.savez_compressed(os.path.join(HERE, "bees_rgb.npz"), rgb=rgb)
        else:
            rgb = gen()
        ok, jpeg, trace, counters, secs = reflib.process_rgb(rgb, q)
        out[name] = {
            "quality": q, "shape": list(rgb.shape), "input_sha256": synth.sha256(rgb), "ok": ok,
            "jpeg_sha256": hashlib.sha256(jpeg).hexdigest(), "jpeg_size": len(jpeg),
            "trace_sha256": hashlib.sha256(trace.encode()).hexdigest(),
            "iterations": counters, "ref_seconds_here": round(secs, 3),
        }
        print(name, out[name]["jpeg_size"], counters, round(secs, 2))
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
