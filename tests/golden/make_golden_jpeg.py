#!/usr/bin/env python3
"""JPEG-input fixtures (row f2): small JPEG files written with Pillow (libjpeg) under
tests/golden/jpeg/ and what the UNMODIFIED reference makes of each of them
(guetzli::Process(jpeg bytes), oracle/_ref) -> tests/golden/golden_jpeg.json.
Runs only where /root/reference exists; files and answers are committed."""
import hashlib
import io
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import reflib  # noqa: E402
from guetzli_b200 import synth  # noqa: E402


def jpeg(rgb, mode="RGB", **kw):
    b = io.BytesIO()
    im = Image.fromarray(rgb)
    if mode != "RGB":
        im = im.convert(mode)
    im.save(b, "JPEG", **kw)
    return b.getvalue()


def exif_blob():
    ex = Image.Exif()
    ex[0x010E] = "guetzli_b200 fixture"  # ImageDescription
    return ex.tobytes()


def cases():
    a = synth.gradnoise(72, 100, 3)
    n = synth.noise(48, 56, 8)
    o = synth.gradnoise(50, 67, 5)
    t = synth.gradnoise(24, 40, 6)
    c = {}
    # (file bytes, quality, clear_metadata)
    c["base444_q90"] = (jpeg(a, quality=90, subsampling=0), 92, True)
    c["prog444_q85"] = (jpeg(a, quality=85, subsampling=0, progressive=True), 92, True)
    c["opt444_q97"] = (jpeg(a, quality=97, subsampling=0, optimize=True), 95, True)
    c["noise444_q92"] = (jpeg(n, quality=92, subsampling=0), 90, True)
    c["odd444_prog"] = (jpeg(o, quality=88, subsampling=0, progressive=True, optimize=True), 88, True)
    c["restart444"] = (jpeg(a, quality=90, subsampling=0, restart_marker_blocks=5), 90, True)
    c["restart444_prog"] = (jpeg(o, quality=90, subsampling=0, progressive=True, restart_marker_rows=1), 90, True)
    c["meta_kept"] = (jpeg(a, quality=90, subsampling=0, exif=exif_blob(), comment=b"hello fixture") + b"TAILBYTES",
                      92, False)
    c["meta_stripped"] = (c["meta_kept"][0], 92, True)
    c["tiny444"] = (jpeg(t, quality=90, subsampling=0), 95, True)
    c["tiny444_meta"] = (jpeg(t, quality=90, subsampling=0, comment=b"tiny"), 95, False)
    c["q100_tables1"] = (jpeg(n, quality=100, subsampling=0), 95, True)
    # rejections
    c["gray"] = (jpeg(a, mode="L", quality=90), 92, True)
    c["sub420"] = (jpeg(a, quality=90, subsampling=2), 92, True)
    c["sub422"] = (jpeg(a, quality=90, subsampling=1), 92, True)
    c["cmyk"] = (jpeg(a, mode="CMYK", quality=90), 92, True)
    c["lowq"] = (c["base444_q90"][0], 80, True)
    c["truncated"] = (c["base444_q90"][0][:1500], 92, True)
    c["garbage"] = (b"\xff\xd8\xff\xe0 this is not a jpeg at all", 92, True)
    bad = bytearray(c["prog444_q85"][0])
    for i in range(600, 640):
        bad[i] = 0x55
    c["corrupt_header"] = (bytes(bad), 92, True)
    bad = bytearray(c["base444_q90"][0])
    for i in range(2000, 2012):
        bad[i] ^= 0x5a
    c["corrupt_scan"] = (bytes(bad), 92, True)
    bad = bytearray(c["prog444_q85"][0])
    for i in range(1800, 1806):
        bad[i] ^= 0x33
    c["corrupt_scan_prog"] = (bytes(bad), 92, True)
    return c


def main():
    out = {}
    os.makedirs(os.path.join(HERE, "jpeg"), exist_ok=True)
    for name, (data, q, clear) in cases().items():
        with open(os.path.join(HERE, "jpeg", name + ".jpg"), "wb") as f:
            f.write(data)
        ok, jpg, trace, counters = reflib.process_jpeg(data, q, clear_metadata=clear)
        rok, dims, coeffs = reflib.read_jpeg(data)
        out[name] = {
            "quality": q, "clear_metadata": clear, "input_sha256": hashlib.sha256(data).hexdigest(), "ok": ok,
            "jpeg_sha256": hashlib.sha256(jpg).hexdigest(), "jpeg_size": len(jpg),
            "trace_sha256": hashlib.sha256(trace.encode()).hexdigest(), "iterations": counters,
            "read_ok": rok, "dims": dims if rok else None,
            "coeffs_sha256": hashlib.sha256(coeffs.tobytes()).hexdigest() if rok else None,
        }
        print(name, len(data), ok, len(jpg), counters, rok)
    json.dump(out, open(os.path.join(HERE, "golden_jpeg.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
