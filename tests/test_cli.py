"""Drop-in `guetzli` CLI (guetzli_b200/cli): its PNG reader against PIL on CPU, and
the reference's smoke-test matrix (tests/smoke_test.sh:39-57) on the GPU."""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest
from PIL import Image

import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PNG_DUMP = os.path.join(ROOT, "oracle", "_build", "png_dump")
CLI = os.path.join(ROOT, "guetzli_b200", "guetzli")
HERE = os.path.dirname(os.path.abspath(__file__))


def blend_on_black(rgb, a):
    return ((rgb.astype(np.int32) * a.astype(np.int32)[..., None] + 128) // 255).astype(np.uint8)


def expected_rgb(path):
    im = Image.open(path)
    if im.mode in ("RGBA", "LA", "PA") or (im.mode == "P" and "transparency" in im.info) or \
            (im.mode in ("L", "RGB", "I;16", "I") and "transparency" in im.info):
        rgba = np.array(im.convert("RGBA"))
        return blend_on_black(rgba[..., :3], rgba[..., 3])
    if im.mode in ("I;16", "I"):
        g = (np.array(im).astype(np.uint32) >> 8).astype(np.uint8)
        return np.stack([g, g, g], axis=-1)
    return np.array(im.convert("RGB"))


def dump(path):
    out = subprocess.run([PNG_DUMP, path], stdout=subprocess.PIPE, check=True).stdout
    w, h = struct.unpack("<ii", out[:8])
    return np.frombuffer(out[8:], dtype=np.uint8).reshape(h, w, 3)


def test_png_reader_matches_pil(tmp_path, port_lib):
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    a = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    cases = {
        "rgb": Image.fromarray(rgb),
        "rgba": Image.fromarray(np.dstack([rgb, a])),
        "gray": Image.fromarray(rgb[..., 0]),
        "la": Image.fromarray(np.dstack([rgb[..., 0], a]), "LA"),
        "pal": Image.fromarray(rgb).quantize(200),
        "pal16": Image.fromarray(rgb).quantize(13),
        "bilevel": Image.fromarray((rgb[..., 0] > 127).astype(np.uint8) * 255).convert("1"),
        "gray16": Image.fromarray((rgb[..., 0].astype(np.uint16) * 257 + 3)),
    }
    for name, im in cases.items():
        p = str(tmp_path / f"{name}.png")
        if name == "pal16":
            im.save(p, bits=4)
        else:
            im.save(p)
        assert np.array_equal(dump(p), expected_rgb(p)), name
    # palette with transparency
    p = str(tmp_path / "palt.png")
    cases["pal"].save(p, transparency=bytes(range(200)))
    assert np.array_equal(dump(p), expected_rgb(p))
    assert subprocess.run([PNG_DUMP, os.devnull]).returncode != 0


def _png_chunk(kind, body):
    import zlib
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xffffffff)


def test_png_reader_rejects_hostile_files(tmp_path, port_lib):
    """The hand-written PNG reader must fail cleanly, before any image-sized allocation, on a
    header that promises more than the data delivers, on data that inflates past what the
    header promises, on oversized dimensions and on a corrupted critical chunk."""
    import zlib
    magic = b"\x89PNG\r\n\x1a\n"

    def png(w, h, raw, ctype=2, depth=8):
        ihdr = struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)
        return magic + _png_chunk(b"IHDR", ihdr) + _png_chunk(b"IDAT", zlib.compress(raw)) + _png_chunk(b"IEND", b"")

    row = lambda w: b"\x00" + bytes(3 * w)
    cases = {
        "ok": (png(4, 3, row(4) * 3), True),
        "huge_header_tiny_data": (png(60000, 60000, row(4) * 3), False),   # would be 10.8 GB of samples
        "zip_bomb": (png(4, 3, bytes(50 * 1024 * 1024)), False),           # inflates far past 39 bytes
        "too_wide": (png(70000, 1, b"\x00" + bytes(3 * 70000)), False),
        "short_data": (png(4, 3, row(4) * 2), False),
    }
    good = png(4, 3, row(4) * 3)
    bad_crc = bytearray(good)
    bad_crc[8 + 8 + 3] ^= 1  # a byte of IHDR's body: CRC no longer matches
    cases["bad_crc"] = (bytes(bad_crc), False)
    for name, (data, ok) in cases.items():
        p = tmp_path / f"{name}.png"
        p.write_bytes(data)
        r = subprocess.run([PNG_DUMP, str(p)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
        assert (r.returncode == 0) == ok, name
    # palette index beyond PLTE: libpng keeps a zero-filled 256-entry palette, i.e. black
    ihdr = struct.pack(">IIBBBBB", 2, 1, 8, 3, 0, 0, 0)
    pal = magic + _png_chunk(b"IHDR", ihdr) + _png_chunk(b"PLTE", bytes([10, 20, 30])) + \
        _png_chunk(b"IDAT", zlib.compress(b"\x00\x00\x05")) + _png_chunk(b"IEND", b"")
    p = tmp_path / "pal_oob.png"
    p.write_bytes(pal)
    assert np.array_equal(dump(str(p)), np.array([[[10, 20, 30], [0, 0, 0]]], dtype=np.uint8))


@pytest.mark.gpu
def test_cli_smoke_matrix(tmp_path, cuda_lib):
    bees = parity.golden_input("bees_444x258_q95")
    png = str(tmp_path / "bees.png")
    Image.fromarray(bees).save(png)
    g = parity.GOLDEN["bees_444x258_q95"]
    out = str(tmp_path / "out.jpg")
    r = subprocess.run([CLI, png, out], stderr=subprocess.PIPE)
    assert r.returncode == 0
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == g["jpeg_sha256"]
    # --verbose trace on stderr, stdin/stdout with "-"
    r = subprocess.run([CLI, "--verbose", "-", "-"], stdin=open(png, "rb"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0
    assert hashlib.sha256(r.stdout).hexdigest() == g["jpeg_sha256"]
    assert hashlib.sha256(r.stderr).hexdigest() == g["trace_sha256"]
    # flags of tests/smoke_test.sh
    small = str(tmp_path / "small.png")
    Image.fromarray(bees[:64, :96]).save(small)
    for flags in (["--nomemlimit"], ["--memlimit", "100"], ["--quality", "85"]):
        r = subprocess.run([CLI] + flags + [small, out], stderr=subprocess.PIPE)
        assert r.returncode == 0 and open(out, "rb").read()[:2] == b"\xff\xd8", flags
    # jpeg input, file and stdin (tests/smoke_test.sh "run_test jpeg ..."): fixture + reference answer
    import json
    gj = json.load(open(os.path.join(HERE, "golden", "golden_jpeg.json")))["base444_q90"]
    jpg_in = os.path.join(HERE, "golden", "jpeg", "base444_q90.jpg")
    r = subprocess.run([CLI, "--quality", str(gj["quality"]), jpg_in, out], stderr=subprocess.PIPE)
    assert r.returncode == 0
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == gj["jpeg_sha256"]
    r = subprocess.run([CLI, "--quality", str(gj["quality"]), "--verbose", "-", "-"], stdin=open(jpg_in, "rb"),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and hashlib.sha256(r.stdout).hexdigest() == gj["jpeg_sha256"]
    assert hashlib.sha256(r.stderr).hexdigest() == gj["trace_sha256"]
    assert subprocess.run([CLI, "--memlimit", "50", jpg_in, out], stderr=subprocess.PIPE).returncode == 1
    # failures: exit code 1
    assert subprocess.run([CLI, os.devnull, out], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([CLI, "--memlimit", "50", small, out], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([CLI, "--quality", "50", small, out], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([CLI, "--bogus", small, out], stderr=subprocess.PIPE).returncode == 1
    assert subprocess.run([CLI, small], stderr=subprocess.PIPE).returncode == 1
