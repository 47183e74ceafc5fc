"""Deterministic synthetic volumes of the shapes BASELINE.json names.

Pure integer hashing (splitmix64) -> floats, so the same seed gives the same
tensors on every host, numpy version and device: golden fixtures, parity tests
and bench.py all draw from here.  Shapes and recipe follow SURVEY.md section 8(d):
LayerNorm'd Gaussian-like features in channels-last layout (what DINOv2's
``x_norm_patchtokens`` gives, models/grid_proto_fewshot.py:90-95), queries derived
from the support slice so matched maps contain foreground, support masks = unions
of ellipses at image resolution, nearest-downsampled like
models/grid_proto_fewshot.py:228-231.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, shape) -> np.ndarray:
    """float32 in [0,1) with 24 random bits, a pure function of (seed, index)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x1000003D1)
    bits = _splitmix64(idx) >> np.uint64(40)
    return (bits.astype(np.float32) * np.float32(2.0 ** -24)).reshape(shape)


def gaussian_like(seed: int, shape) -> np.ndarray:
    """Sum of four uniforms, centred and scaled to unit variance (Irwin-Hall)."""
    u = sum(uniform(seed * 4 + k, shape) for k in range(4))
    return ((u - np.float32(2.0)) * np.float32(np.sqrt(3.0))).astype(np.float32)


def layer_norm(x: np.ndarray) -> np.ndarray:
    m = x.mean(-1, keepdims=True, dtype=np.float32)
    v = ((x - m) ** 2).mean(-1, keepdims=True, dtype=np.float32)
    return ((x - m) / np.sqrt(v + np.float32(1e-6))).astype(np.float32)


def ellipse_mask(seed: int, size: int, n_ell: int | None = None, lo=0.03, hi=0.25) -> np.ndarray:
    """Union of 1-3 axis-aligned ellipses covering roughly lo..hi of a size x size image."""
    r = uniform(seed, (16,))
    k = n_ell or 1 + int(r[0] * 3)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    m = np.zeros((size, size), bool)
    target = lo + (hi - lo) * float(r[1])
    for e in range(k):
        cx = (0.25 + 0.5 * float(r[2 + 4 * e])) * size
        cy = (0.25 + 0.5 * float(r[3 + 4 * e])) * size
        area = target * size * size / k
        aspect = 0.6 + 0.8 * float(r[4 + 4 * e])
        a = np.sqrt(area / np.pi * aspect)
        b = area / (np.pi * a)
        m |= ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0
    return m.astype(np.float32)


def nearest_resize(mask: np.ndarray, h: int, w: int) -> np.ndarray:
    """F.interpolate(mask, (h,w), mode='nearest'): src = floor(dst * in / out)."""
    H, W = mask.shape[-2:]
    ys = np.minimum((np.arange(h, dtype=np.float32) * np.float32(H / h)).astype(np.int64), H - 1)
    xs = np.minimum((np.arange(w, dtype=np.float32) * np.float32(W / w)).astype(np.int64), W - 1)
    return mask[..., ys[:, None], xs[None, :]]


@dataclass
class Volume:
    """One synthetic volume.  Features are channels-last: sup [S,h,w,C], qry [Q,h,w,C]."""
    sup: np.ndarray          # [S,h,w,C] float32
    qry: np.ndarray          # [Q,h,w,C] float32
    fg_img: np.ndarray       # [L,S,img,img] float32 {0,1} support masks at image resolution
    fg: np.ndarray           # [L,S,h,w] nearest-downsampled foreground masks
    img_size: int

    @property
    def bg(self) -> np.ndarray:
        return (1.0 - self.fg).astype(np.float32)     # ProtoSAM.py:63


def make_volume(seed: int, Q: int, L: int, C: int, h: int, w: int, img_size: int, S: int = 1,
                noise: float = 0.3) -> Volume:
    sup = layer_norm(gaussian_like(seed * 7 + 1, (S, h, w, C)))
    # smooth the support features a little in space so coarse maps have blobs, not salt-and-pepper
    sup = layer_norm((sup + np.roll(sup, 1, 1) + np.roll(sup, 1, 2) + np.roll(sup, (1, 1), (1, 2))) / 4.0)
    qry = np.empty((Q, h, w, C), np.float32)
    for q in range(Q):
        shift = (q % 5) - 2
        base = np.roll(sup[q % S], (shift, -shift), (0, 1))
        qry[q] = base + np.float32(noise) * gaussian_like(seed * 7 + 1000 + q, (h, w, C))
    fg_img = np.stack([np.stack([ellipse_mask(seed * 131 + l * 17 + s, img_size) for s in range(S)])
                       for l in range(L)])
    fg = nearest_resize(fg_img, h, w).astype(np.float32)
    return Volume(sup=sup, qry=qry, fg_img=fg_img, fg=fg, img_size=img_size)


# the named configs of BASELINE.json; Q may be overridden for tests
CONFIGS = {
    "cfg1_vits_256": dict(Q=1, L=1, C=384, h=32, w=32, img_size=256, ws=2),
    "cfg2_chaos_mri": dict(Q=32, L=4, C=768, h=37, w=37, img_size=518, ws=2),
    "cfg3_synapse_ct": dict(Q=128, L=4, C=1024, h=48, w=48, img_size=672, ws=2),
    "cfg4_polyp_1024": dict(Q=64, L=1, C=1024, h=73, w=73, img_size=1024, ws=2),
    "cfg5_stress_vitl": dict(Q=1024, L=1, C=1024, h=48, w=48, img_size=672, ws=2),
}
