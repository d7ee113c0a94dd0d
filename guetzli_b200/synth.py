"""Synthetic sRGB inputs named by BASELINE.json / SURVEY.md §8(d).

numpy >= 2, PCG64 via np.random.default_rng(seed); every caller should log
sha256 of the returned bytes.
"""
import hashlib
import numpy as np


def noise(h, w, seed):
    """Uniform sRGB noise: rng.integers(0, 256, (H, W, 3), uint8)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def gradnoise(h, w, seed, sigma=8.0):
    """Colour gradient plus Gaussian noise (sigma in 8-bit code values)."""
    rng = np.random.default_rng(seed)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    img = np.empty((h, w, 3), dtype=np.float64)
    img[..., 0] = x * 255.0 / (w - 1) + 0 * y
    img[..., 1] = y * 255.0 / (h - 1) + 0 * x
    img[..., 2] = (x + y) * 255.0 / (w + h - 2)
    img += rng.normal(0.0, sigma, (h, w, 3))
    return np.round(np.clip(img, 0.0, 255.0)).astype(np.uint8)


def sha256(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
