"""Procedural YUV clips (integer-only, counter-based, deterministic) for tests and bench.

The BASELINE configs are far larger than host RAM, and the reference ships no
clips (``.gitignore:56`` excludes ``*.mp4``), so every workload is synthesised: two moving sinusoid
gratings + a slow pan + hash noise, with a flat patch (VIF non-log branch), and a distorted twin =
3-tap box blur + quantisation + noise, with a sharpened patch (gain > 1, where the NEG models
differ) and an inverted patch (sigma12 < 0)."""
from __future__ import annotations

import numpy as np

_SIN = np.round(np.sin(np.arange(1024) * (2 * np.pi / 1024)) * 1024).astype(np.int32)


def _hash(seed: int, f: int, h: int, w: int) -> np.ndarray:
    y = np.arange(h, dtype=np.uint32)[:, None]
    x = np.arange(w, dtype=np.uint32)[None, :]
    k = np.uint32((seed * 0xC2B2AE3D + f * 0x27D4EB2F + 0x165667B1) & 0xFFFFFFFF)
    v = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ k
    v ^= v >> np.uint32(15)
    v *= np.uint32(0x2C1B3C6D)
    v ^= v >> np.uint32(12)
    v *= np.uint32(0x297A2D39)
    v ^= v >> np.uint32(15)
    return v


def ref_luma(seed: int, f: int, w: int, h: int, bpc: int = 8) -> np.ndarray:
    """Reference luma plane of frame f: values roughly uniform over the legal range."""
    y = np.arange(h, dtype=np.int32)[:, None]
    x = np.arange(w, dtype=np.int32)[None, :]
    g1 = _SIN[(3 * x + 2 * y + 5 * f + 131 * seed) & 1023] * 52 >> 10
    g2 = _SIN[(23 * x - 17 * y + 11 * f) & 1023] * 22 >> 10
    ramp = ((x + 2 * f) * 64 // max(w, 1)) - 32
    n = (_hash(seed, f, h, w) & np.uint32(15)).astype(np.int32) - 8
    v = 126 + g1 + g2 + ramp + n
    # flat patch: tiny variance -> sigma1_sq < sigma_nsq
    y0, y1, x0, x1 = h // 8, h // 4, w // 8, w // 3
    v[y0:y1, x0:x1] = 90 + (n[y0:y1, x0:x1] & 1)
    v = np.clip(v, 16, 235)
    if bpc == 8:
        return v.astype(np.uint8)
    extra = ((_hash(seed + 7, f, h, w) >> np.uint32(9)) & np.uint32((1 << (bpc - 8)) - 1)).astype(np.int32)
    return ((v << (bpc - 8)) + extra).astype(np.uint16)


def _box3(a: np.ndarray) -> np.ndarray:
    p = np.pad(a.astype(np.int32), 1, mode="edge")
    hsum = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    return (hsum[:-2] + 2 * hsum[1:-1] + hsum[2:] + 8) >> 4


def distort(ref: np.ndarray, seed: int, f: int, bpc: int = 8, q: int = 4, strength: int = 1) -> np.ndarray:
    """Distorted twin of a plane (works for luma and chroma)."""
    h, w = ref.shape
    sc = 1 << (bpc - 8)
    r = ref.astype(np.int32)
    b = _box3(ref)
    for _ in range(strength - 1):
        b = _box3(b)
    qq = q * sc
    v = (b // qq) * qq + qq // 2
    n = (_hash(seed + 1000, f, h, w) & np.uint32(7)).astype(np.int32) - 4
    v = v + n * sc // 2
    # sharpened patch (enhancement: g > 1)
    y0, y1, x0, x1 = h // 2, h // 2 + h // 5, w // 2, w // 2 + w // 4
    v[y0:y1, x0:x1] = (3 * r[y0:y1, x0:x1] - 2 * b[y0:y1, x0:x1])
    # inverted patch (sigma12 < 0)
    y0, y1, x0, x1 = h // 3, h // 3 + h // 10, w // 10, w // 10 + w // 6
    v[y0:y1, x0:x1] = (255 * sc) - r[y0:y1, x0:x1]
    v = np.clip(v, 0, (1 << bpc) - 1)
    return v.astype(ref.dtype)


def ref_chroma(seed: int, f: int, w: int, h: int, bpc: int, which: int) -> np.ndarray:
    y = np.arange(h, dtype=np.int32)[:, None]
    x = np.arange(w, dtype=np.int32)[None, :]
    g = _SIN[(5 * x + (3 + which) * y + 7 * f + 977 * which) & 1023] * 30 >> 10
    n = (_hash(seed + 31 + which, f, h, w) & np.uint32(7)).astype(np.int32) - 4
    v = np.clip(128 + g + n, 16, 240)
    if bpc == 8:
        return v.astype(np.uint8)
    return (v << (bpc - 8)).astype(np.uint16)


def frame_pair(seed: int, f: int, w: int, h: int, bpc: int = 8, chroma: bool = True, q: int = 4,
               strength: int = 1):
    """Returns (ref_planes, dis_planes): lists [Y, U, V] (or [Y]) of C-contiguous arrays (4:2:0)."""
    ry = ref_luma(seed, f, w, h, bpc)
    dy = distort(ry, seed, f, bpc, q, strength)
    if not chroma:
        return [ry], [dy]
    cw, ch = (w + 1) // 2, (h + 1) // 2
    rp, dp = [ry], [dy]
    for k in (1, 2):
        c = ref_chroma(seed, f, cw, ch, bpc, k)
        rp.append(c)
        dp.append(distort(c, seed + k, f, bpc, q, strength))
    return rp, dp
