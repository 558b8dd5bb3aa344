"""Seeded synthetic "EuRoC-shaped" frames and LightGlue inputs.

TEST/BENCH INFRASTRUCTURE: pure numpy so the same bytes are produced on every box.
"""
from __future__ import annotations

import numpy as np


def _gauss_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(4 * sigma + 0.5)
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    pad = np.pad(img, ((r, r), (r, r)), mode="reflect")
    tmp = np.zeros_like(img)
    for i, w in enumerate(k):
        tmp += w * pad[i:i + img.shape[0], r:r + img.shape[1]]
    pad = np.pad(tmp, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * pad[:, i:i + img.shape[1]]
    return out


def canvas(seed: int, h: int, w: int) -> np.ndarray:
    """u8 [h, w]: blurred noise + 40 grey rectangles (about 1.6 k SuperPoint keypoints at 640x480)."""
    rng = np.random.RandomState(seed)
    img = rng.rand(h, w) * 255.0
    img = _gauss_blur(img, 2.0)
    img = (img - img.min()) / (img.max() - img.min()) * 255.0
    n_rect = max(1, int(round(40 * (h * w) / (480.0 * 640.0))))
    for _ in range(n_rect):
        x0 = rng.randint(0, max(1, w - 40))
        y0 = rng.randint(0, max(1, h - 40))
        rw = rng.randint(10, 80)
        rh = rng.randint(10, 80)
        g = rng.randint(0, 255)
        img[y0:y0 + rh, x0:x0 + rw] = g
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frame(seed: int, h: int = 480, w: int = 640) -> np.ndarray:
    return canvas(seed, h, w)


def frame_pair(seed: int, h: int = 480, w: int = 640, shift=(12, 7), margin: int = 20):
    """Two views of one canvas; view B is view A's content translated by shift=(dx, dy) px."""
    dx, dy = shift
    big = canvas(seed, h + 2 * margin, w + 2 * margin)
    a = big[margin:margin + h, margin:margin + w]
    b = big[margin - dy:margin - dy + h, margin - dx:margin - dx + w]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def lightglue_inputs(n: int, seed: int, h: int = 480, w: int = 640, noise: float = 0.05):
    """random integer keypoints, unit descriptors, permuted+noised second set."""
    rng = np.random.RandomState(seed)
    cells = rng.permutation((w - 8) * (h - 8))[:n]
    k0 = np.stack([cells % (w - 8) + 4, cells // (w - 8) + 4], 1).astype(np.float32)
    d0 = rng.randn(n, 256).astype(np.float32)
    d0 /= np.linalg.norm(d0, axis=1, keepdims=True)
    perm = rng.permutation(n)
    k1 = k0[perm] + rng.randint(-2, 3, size=(n, 2)).astype(np.float32)
    d1 = d0[perm] + noise * rng.randn(n, 256).astype(np.float32)
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    return k0, k1.astype(np.float32), d0, d1.astype(np.float32), perm
