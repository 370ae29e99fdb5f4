"""Seeded synthetic inputs for the SSD multibox head path.

numpy's PCG64 stream is used (not torch's CPU generator) so that the same seed
gives the same bits in this container and on the GPU box: the golden fixtures
under ``tests/golden/`` store a digest of the inputs they were made from.

Shapes follow the reference's input contract (``Dataset.py:26,36``,
``Model.py:235``): per image a float32 class vector ``[n_i]`` with ids 0..19
and fractional xyxy boxes ``[n_i, 4]``; head outputs ``loc [B,P,4]`` and
``conf [B,P,C]`` in float32.
"""
from __future__ import annotations

import hashlib

import numpy as np


def make_gt(seed: int, batch: int, min_gt: int = 1, max_gt: int = 10):
    """Ragged ground truth: centres U(0.1,0.9)^2, sizes U(0.05,0.45)^2, clamped to [0,1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    boxes, classes = [], []
    for _ in range(batch):
        n = int(rng.integers(min_gt, max_gt + 1))
        c = rng.uniform(0.1, 0.9, size=(n, 2)).astype(np.float32)
        s = rng.uniform(0.05, 0.45, size=(n, 2)).astype(np.float32)
        half = s / np.float32(2)
        b = np.concatenate([c - half, c + half], axis=1).astype(np.float32)
        np.clip(b, 0.0, 1.0, out=b)
        boxes.append(b)
        classes.append(rng.integers(0, 20, size=(n,)).astype(np.float32))
    return boxes, classes


def make_head(seed: int, batch: int, num_priors: int, num_classes: int = 21,
              loc_scale: float = 1.0, bg_bias: float = 0.0):
    """Head outputs: loc ~ loc_scale*N(0,1), conf ~ N(0,1) with the background
    logit (last class, ``Losses.py:171``) shifted by ``bg_bias``."""
    rng = np.random.Generator(np.random.PCG64(seed + 1_000_003))
    loc = rng.standard_normal((batch, num_priors, 4), dtype=np.float32)
    if loc_scale != 1.0:
        loc *= np.float32(loc_scale)
    conf = rng.standard_normal((batch, num_priors, num_classes), dtype=np.float32)
    if bg_bias != 0.0:
        conf[:, :, num_classes - 1] += np.float32(bg_bias)
    return loc, conf


def pack_gt(boxes, classes):
    """Ragged lists -> packed ``[sum G,4]`` boxes, ``[sum G]`` classes, ``[B+1]`` int32 offsets
    (the host-side equivalent of ``Losses.py:129-130``)."""
    counts = [int(b.shape[0]) for b in boxes]
    off = np.zeros(len(boxes) + 1, dtype=np.int32)
    off[1:] = np.cumsum(counts)
    if off[-1] == 0:
        return np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), off
    gb = np.concatenate([np.asarray(b, np.float32).reshape(-1, 4) for b in boxes], 0)
    gc = np.concatenate([np.asarray(c, np.float32).reshape(-1) for c in classes], 0)
    return np.ascontiguousarray(gb), np.ascontiguousarray(gc), off


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()[:32]
