"""gt collate: the step before ``ssd()``.

The reference builds a batch's ground truth as two Python lists of small tensors (``Dataset.py:24-53``), moves them to
the device one by one (``train_function.py:62-63``: 2B tiny copies) and concatenates them again inside the loss
(``Losses.py:129-130``).  ``collate_gt`` does the difficult-filter (``Dataset.py:28-30``), the division by the image
size (``Dataset.py:35-36``) and the packing in ONE native host pass (``ssdhead_pack_gt``) into page-locked buffers, so
the batch's gt reaches the device with a single asynchronous copy per array in the layout every entry point takes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib


def _as_f32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def collate_gt(boxes: Sequence, classes: Sequence, difficult: Optional[Sequence] = None, keep_difficult: bool = True,
               img_wh=None, pinned: bool = True, out=None):
    """Ragged per-image gt -> (boxes [sumG,4] f32, classes [sumG] f32, offsets int32 [B+1]) as numpy arrays.

    ``boxes[i]`` is ``[n_i,4]`` xyxy (pixels when ``img_wh`` [B,2] is given, else already fractional), ``classes[i]``
    ``[n_i]`` class ids, ``difficult[i]`` ``[n_i]`` flags.  With ``pinned`` the arrays live in page-locked memory
    (``ssdhead_host_alloc``) when a CUDA device is present; every returned array keeps that allocation alive on its
    own, so the result may be unpacked, sliced and the tuple dropped.  ``out``: a ``pinned.Staging`` to pack into
    (the per-step upload path of ``ssd()`` recycles its blocks).  Raises ``IndexError`` for an image left without a
    box, as the reference does (``Losses.py:153``)."""
    lib = _lib.load()
    B = len(boxes)
    if len(classes) != B or (difficult is not None and len(difficult) != B):
        raise ValueError("collate_gt: boxes, classes and difficult must list the same images")
    bx = [_as_f32(b).reshape(-1, 4) for b in boxes]
    cl = [_as_f32(c).reshape(-1) for c in classes]
    counts = np.array([b.shape[0] for b in bx], dtype=np.int32)
    for i in range(B):
        if cl[i].shape[0] != counts[i]:
            raise ValueError(f"collate_gt: image {i} has {counts[i]} boxes but {cl[i].shape[0]} classes")
    df = None
    if difficult is not None:
        df = [np.ascontiguousarray(d.detach().cpu().numpy() if hasattr(d, "detach") else d).astype(np.uint8).reshape(-1)
              for d in difficult]
    cap = int(counts.sum())
    wh = None if img_wh is None else _as_f32(img_wh).reshape(B, 2)

    def ptr_array(arrs):
        return (C.c_void_p * max(B, 1))(*[a.ctypes.data if a.size else None for a in arrs])

    if out is not None:
        if out.capacity < cap or out.B != B:
            raise ValueError("collate_gt: staging block too small for this batch")
        out_b, out_c, out_o = out.boxes, out.classes, out.offsets
        cap = out.capacity
    else:
        out_b, out_c, out_o = _alloc(cap, B, pinned)
    rc = lib.ssdhead_pack_gt(ptr_array(bx), ptr_array(cl), ptr_array(df) if df is not None else None,
                             counts.ctypes.data, B, 1 if keep_difficult else 0,
                             wh.ctypes.data if wh is not None else None,
                             out_b.ctypes.data, out_c.ctypes.data, out_o.ctypes.data, cap)
    if rc == _lib.E_STATE:
        raise IndexError("collate_gt: an image has no ground-truth box left: max() over an empty dimension")
    if rc < 0:
        _lib.check(rc, "ssdhead_pack_gt")
    n = int(rc)
    return out_b[:n], out_c[:n], out_o


def _alloc(cap, B, pinned):
    """Three arrays for `cap` boxes; page-locked (one block, each array owning it) when a CUDA device is present."""
    if pinned:
        try:
            from .pinned import Staging
            s = Staging(cap, B)
            return s.boxes, s.classes, s.offsets
        except RuntimeError:
            pass                                  # no CUDA device / no page-locked memory: ordinary host arrays
    return np.empty((cap, 4), np.float32), np.empty((cap,), np.float32), np.empty((B + 1,), np.int32)
