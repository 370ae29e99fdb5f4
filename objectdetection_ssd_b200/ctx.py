"""Python handle of ``ssdhead_ctx`` (include/ssdhead.h): one C call per step.

``SSDHeadContext.loss_host`` / ``detect_host`` take HOST arrays (numpy, ideally page-locked via
``pinned_empty``) - the same buffers the reference's CPU path works on - and run the whole step
(copies in, kernels, copies out, pipelined in image chunks) inside the library.
``loss_dev`` runs one training-head step on device tensors with the match overlapped with the
CE streaming kernel.  No torch arithmetic is involved; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


from .pinned import pinned_empty       # noqa: E402,F401  (page-locked host array that owns its allocation)


def pinned_free(arr: np.ndarray) -> None:
    """Kept for callers of the first ABI revision: page-locked arrays now release their allocation when the last view of
    them is collected, so there is nothing to do here."""
    return None


class SparseRows:
    """Output buffers of ``loss_host_sparse`` (layout: ``ssdhead_mine_sparse`` in include/ssdhead.h), page-locked by
    default so the mining kernel writes its rows straight into them.  ``cnt[b] = (rows, positives)`` of image b;
    ``idx[b, s]`` = prior of slot s; ``grad_conf[b, s]`` its conf-gradient row; ``grad_loc[b, s]`` for s < positives."""

    def __init__(self, B: int, cap: int = 1024, C_: int = 21, pinned: bool = True):
        alloc = pinned_empty if pinned else (lambda shape, dt=np.float32: np.empty(shape, dt))
        self.B, self.cap, self.C = int(B), int(cap), int(C_)
        self.cnt = alloc((B, 2), np.int32)
        self.idx = alloc((B, cap), np.int32)
        self.grad_conf = alloc((B, cap, C_), np.float32)
        self.grad_loc = alloc((B, cap, 4), np.float32)
        self.cnt[...] = 0

    def scatter(self, P: int, B: int = None):
        """Dense (grad_loc [B,P,4], grad_conf [B,P,C]) numpy arrays rebuilt from the rows (tests / small batches)."""
        B = self.B if B is None else B
        gl = np.zeros((B, P, 4), np.float32)
        gc = np.zeros((B, P, self.C), np.float32)
        for b in range(B):
            n, npos = int(self.cnt[b, 0]), int(self.cnt[b, 1])
            if n > self.cap:
                raise RuntimeError(f"image {b} produced {n} gradient rows, the buffers hold {self.cap}")
            gc[b, self.idx[b, :n]] = self.grad_conf[b, :n]
            gl[b, self.idx[b, :npos]] = self.grad_loc[b, :npos]
        return gl, gc

    def nbytes_used(self, B: int = None) -> int:
        """Bytes the device actually wrote (rows + indices + counts) - what crossed PCIe."""
        B = self.B if B is None else B
        n = int(np.minimum(self.cnt[:B, 0], self.cap).sum())
        npos = int(self.cnt[:B, 1].sum())
        return n * (self.C * 4 + 4) + npos * 16 + B * 8


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SSDHeadContext:
    def __init__(self, priors_cxcywh: np.ndarray, max_batch: int, num_classes: int = 21, max_total_gt: int = 0,
                 top_k: int = 200, device: int = 0):
        self.lib = _lib.load()
        pri = np.ascontiguousarray(priors_cxcywh, dtype=np.float32)
        self.P, self.C, self.maxB, self.top_k = int(pri.shape[0]), int(num_classes), int(max_batch), int(top_k)
        self.max_total_gt = int(max_total_gt) if max_total_gt > 0 else 128 * self.maxB
        h = C.c_void_p()
        _lib.check(self.lib.ssdhead_ctx_create(C.byref(h), device, self.maxB, self.P, self.C, self.max_total_gt,
                                               self.top_k, _p(pri)), "ssdhead_ctx_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ssdhead_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host-buffer entry points
    def loss_host(self, loc, conf, gt_xyxy, gt_cls, gt_off, grad_loc=None, grad_conf=None,
                  neg_ratio: int = 3, pos_iou: float = 0.5):
        """ssd() on host arrays.  Returns (loc_loss, conf_loss); gradients are written into
        ``grad_loc`` / ``grad_conf`` when given (pass both or neither)."""
        B = int(loc.shape[0])
        losses = np.zeros(2, np.float32)
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_host(
            self._h, _p(loc), _p(conf), _p(gt_xyxy), _p(gt_cls), _p(gt_off), B, int(neg_ratio), float(pos_iou),
            _p(losses), _p(grad_loc), _p(grad_conf)), "ssdhead_ctx_multibox_loss_host")
        return float(losses[0]), float(losses[1])

    def loss_host_sparse(self, loc, conf, gt_xyxy, gt_cls, gt_off, rows: "SparseRows",
                         neg_ratio: int = 3, pos_iou: float = 0.5):
        """ssd() on host arrays with the gradients returned as packed rows in ``rows`` (``SparseRows``): nothing dense
        is zeroed or copied.  Returns (loc_loss, conf_loss)."""
        B = int(loc.shape[0])
        if B > rows.B:
            raise ValueError(f"SparseRows holds {rows.B} images, the batch has {B}")
        losses = np.zeros(2, np.float32)
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_host_sparse(
            self._h, _p(loc), _p(conf), _p(gt_xyxy), _p(gt_cls), _p(gt_off), B, int(neg_ratio), float(pos_iou),
            _p(losses), rows.cap, _p(rows.cnt), _p(rows.idx), _p(rows.grad_conf), _p(rows.grad_loc)),
            "ssdhead_ctx_multibox_loss_host_sparse")
        return float(losses[0]), float(losses[1])

    def detect_host(self, loc, conf, out_boxes, out_prob, out_cls, out_prior, out_cnt,
                    min_score: float = 0.2, iou_thr: float = 0.45):
        """inference() over a batch of host arrays; outputs as ``ssdhead_detect``."""
        B = int(loc.shape[0])
        _lib.check(self.lib.ssdhead_ctx_detect_host(
            self._h, _p(loc), _p(conf), B, float(min_score), float(iou_thr),
            _p(out_boxes), _p(out_prob), _p(out_cls), _p(out_prior), _p(out_cnt)), "ssdhead_ctx_detect_host")

    # ------------------------------------------------------------------ device-resident step
    def loss_dev(self, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, B, sumG, sums_ptr, losses_ptr,
                 grad_loc_ptr, grad_conf_ptr, stream, neg_ratio: int = 3, pos_iou: float = 0.5):
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_dev(
            self._h, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, int(B), int(sumG), int(neg_ratio),
            float(pos_iou), sums_ptr, losses_ptr, grad_loc_ptr, grad_conf_ptr, stream), "ssdhead_ctx_multibox_loss_dev")

    def loss_dev_resident(self, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, B, sumG, sums_ptr, losses_ptr,
                          grad_loc_ptr, grad_conf_ptr, stream, fresh: bool = False, neg_ratio: int = 3, pos_iou: float = 0.5):
        """``loss_dev`` with RESIDENT gradient tensors: the same ``grad_loc`` / ``grad_conf`` from step to step; every
        call retracts the rows the previous call wrote and writes its own (the gradient is sparse: ~4 Npos of 8732 rows
        per image), so no dense zero background is written.  ``fresh=True`` on the first call, with new tensors, or after
        anybody else wrote them (the context zero-fills them).  Same bits as ``loss_dev``."""
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_dev_resident(
            self._h, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, int(B), int(sumG), int(neg_ratio),
            float(pos_iou), sums_ptr, losses_ptr, grad_loc_ptr, grad_conf_ptr, 1 if fresh else 0, stream),
            "ssdhead_ctx_multibox_loss_dev_resident")

    def loss_levels_dev(self, levels, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, B, sumG, sums_ptr, losses_ptr, stream,
                        neg_ratio: int = 3, pos_iou: float = 0.5):
        """``loss_dev`` on per-level head tensors: ``levels`` is a filled ``_lib.Levels`` (conf / loc / grad pointers per
        level).  After ``xchg_import`` this is the sharded two-kernel step with global normalisation."""
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_levels_dev(
            self._h, C.addressof(levels), gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, int(B), int(sumG), int(neg_ratio),
            float(pos_iou), sums_ptr, losses_ptr, stream), "ssdhead_ctx_multibox_loss_levels_dev")

    def loss_begin(self, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, B, sumG, grad_loc_ptr, grad_conf_ptr, stream,
                   pos_iou: float = 0.5) -> int:
        """First half of a sharded step; returns the device address of this rank's int32 positive count."""
        out = C.c_void_p()
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_begin(
            self._h, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, int(B), int(sumG), float(pos_iou),
            grad_loc_ptr, grad_conf_ptr, C.byref(out), stream), "ssdhead_ctx_multibox_loss_begin")
        return int(out.value)

    def loss_end(self, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, B, npos_norm_ptr, sums_ptr, losses_ptr,
                 grad_loc_ptr, grad_conf_ptr, stream, neg_ratio: int = 3, pos_iou: float = 0.5):
        _lib.check(self.lib.ssdhead_ctx_multibox_loss_end(
            self._h, loc_ptr, conf_ptr, gt_xyxy_ptr, gt_cls_ptr, gt_off_ptr, int(B), int(neg_ratio), float(pos_iou),
            npos_norm_ptr, sums_ptr, losses_ptr, grad_loc_ptr, grad_conf_ptr, stream), "ssdhead_ctx_multibox_loss_end")

    # ------------------------------------------------------------------ sharded batches over NVLink peer memory
    def xchg_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's exchange buffer (gather them across ranks, then ``xchg_import``)."""
        buf = C.create_string_buffer(64)
        _lib.check(self.lib.ssdhead_ctx_xchg_export(self._h, buf), "ssdhead_ctx_xchg_export")
        return buf.raw

    def xchg_import(self, handles, rank: int) -> None:
        """Map every rank's exchange buffer; afterwards ``loss_dev`` runs the sharded two-kernel step (global
        normalisation, no NCCL call).  All ranks must then call ``loss_dev`` in lock step."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        _lib.check(self.lib.ssdhead_ctx_xchg_import(self._h, blob, len(handles), int(rank)), "ssdhead_ctx_xchg_import")

    def xchg_error(self) -> bool:
        return bool(self.lib.ssdhead_ctx_xchg_error(self._h))

    def finish_loss(self, sums_ptr, npos_norm_ptr, losses_ptr, stream):
        _lib.check(self.lib.ssdhead_finish_loss(sums_ptr, npos_norm_ptr, losses_ptr, stream), "ssdhead_finish_loss")
