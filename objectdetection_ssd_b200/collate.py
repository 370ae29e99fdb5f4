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
               img_wh=None, pinned: bool = True):
    """Ragged per-image gt -> (boxes [sumG,4] f32, classes [sumG] f32, offsets int32 [B+1]) as numpy arrays.

    ``boxes[i]`` is ``[n_i,4]`` xyxy (pixels when ``img_wh`` [B,2] is given, else already fractional), ``classes[i]``
    ``[n_i]`` class ids, ``difficult[i]`` ``[n_i]`` flags.  With ``pinned`` the arrays live in page-locked memory
    (``ssdhead_host_alloc``) when a CUDA device is present.  Raises ``IndexError`` for an image left without a box,
    as the reference does (``Losses.py:153``)."""
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

    out_b, out_c, out_o, owner = _alloc(lib, cap, B, pinned)
    rc = lib.ssdhead_pack_gt(ptr_array(bx), ptr_array(cl), ptr_array(df) if df is not None else None,
                             counts.ctypes.data, B, 1 if keep_difficult else 0,
                             wh.ctypes.data if wh is not None else None,
                             out_b.ctypes.data, out_c.ctypes.data, out_o.ctypes.data, cap)
    if rc == _lib.E_STATE:
        raise IndexError("collate_gt: an image has no ground-truth box left: max() over an empty dimension")
    if rc < 0:
        _lib.check(rc, "ssdhead_pack_gt")
    n = int(rc)
    res = (out_b[:n], out_c[:n], out_o)
    for r in res:                      # keep the page-locked allocation alive as long as any view is
        r.flags.writeable = True
    return _Packed(res, owner)


class _Packed(tuple):
    """(boxes, classes, offsets) - a tuple that also owns the page-locked allocation behind the arrays."""

    def __new__(cls, arrays, owner):
        self = super().__new__(cls, arrays)
        self._owner = owner
        return self


class _PinnedBlock:
    def __init__(self, lib, nbytes):
        self.lib = lib
        self.ptr = lib.ssdhead_host_alloc(max(nbytes, 16))

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.ssdhead_host_free(self.ptr)
            self.ptr = None


def _alloc(lib, cap, B, pinned):
    nb, nc, no = cap * 16, cap * 4, (B + 1) * 4
    if pinned:
        blk = _PinnedBlock(lib, nb + nc + no + 64)
        if blk.ptr:
            def view(off, count, dt):
                buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(blk.ptr + off)
                return np.frombuffer(buf, dtype=dt, count=count)
            o1 = (nb + 15) // 16 * 16
            o2 = o1 + (nc + 15) // 16 * 16
            return view(0, cap * 4, np.float32).reshape(cap, 4), view(o1, cap, np.float32), view(o2, B + 1, np.int32), blk
    return np.empty((cap, 4), np.float32), np.empty((cap,), np.float32), np.empty((B + 1,), np.int32), None
