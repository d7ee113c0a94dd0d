#!/usr/bin/env python3
"""Differential run of guetzli::Process(JPEG bytes) -- CPU port of the product's code against the unmodified
reference -- on random 4:4:4 JPEG files written by PIL (baseline / progressive / optimised tables, any
quality, with and without metadata stripping): same verdict, JPEG bytes and --verbose trace.
usage: tools/fuzz_jpeg_input.py [first_seed] [count]"""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

import guetzli_b200 as gb  # noqa: E402
import reflib  # noqa: E402
from fuzz_parity import random_image  # noqa: E402


def one(seed, port):
    rng = np.random.default_rng(seed)
    rgb, kind = random_image(rng)
    buf = io.BytesIO()
    src_q = int(rng.integers(50, 101))
    kw = dict(quality=src_q, subsampling=0, progressive=bool(rng.random() < 0.4), optimize=bool(rng.random() < 0.5))
    if rng.random() < 0.3:
        kw["comment"] = b"fuzz"
    Image.fromarray(rgb).save(buf, "JPEG", **kw)
    data = buf.getvalue()
    quality = float(rng.choice([84, 88, 90, 95, 97, 100]))
    clear = bool(rng.random() < 0.6)
    rok, rjpeg, rtrace, _ = reflib.process_jpeg(data, quality, clear_metadata=clear)
    st = gb.ProcessStats(debug_output=[])
    p = gb.Params(butteraugli_target=gb.butteraugli_score_for_quality(quality, lib=port), clear_metadata=clear)
    ok, jpeg = gb.process_jpeg(p, st, data, lib=port)
    same = ok == rok and jpeg == rjpeg and "".join(st.debug_output) == rtrace
    return same, dict(seed=seed, kind=kind, shape=rgb.shape, src=kw, quality=quality, clear=clear, ok=ok, rok=rok)


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    port = gb.load_library(os.path.join(ROOT, "oracle", "_build", "libguetzli_port.so"))
    bad = 0
    for seed in range(first, first + count):
        same, info = one(seed, port)
        if not same:
            bad += 1
            print("MISMATCH", info, flush=True)
        if (seed - first + 1) % 25 == 0:
            print(f"... {seed - first + 1} cases, {bad} mismatches", flush=True)
    print(f"{count} cases from seed {first}: {bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
