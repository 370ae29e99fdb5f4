"""CPU oracle for the SSD multibox head path.   *** TEST INFRASTRUCTURE ***

A restatement, in torch-CPU / numpy float32 arithmetic, of the algorithm the
reference (nitishsaDire/objectDetection_ssd) runs on this path.  It exists to
CHECK the CUDA kernels; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``objectdetection_ssd_b200``) never does, and has no CPU
fallback.

Parity status: PINNED.  The reference holds no tests or golden vectors
(SURVEY.md section 4), so the pin is the reference itself: every function here is
compared bit-for-bit / to 1e-6 against the unmodified reference imported from
``/root/reference`` (``tests/test_oracle_vs_reference.py``, run where the
reference is mounted) and against fixtures frozen from that reference under
``tests/golden/`` (``tests/golden/make_golden.py`` is the generating script).

Every function cites the reference lines it restates.  All arithmetic is
float32 with the reference's operation order: IoU uses only + - * / min max, so
it is bit-reproducible on any IEEE machine and the integer outputs derived from
it (match indices, class map, positive mask) are exact.  exp/log/softmax go
through the same ATen CPU kernels the reference calls.

Tie rules the oracle fixes where torch leaves them open:
  T1  best gt per prior, equal IoU      -> lowest gt index   (Tensor.max = first maximum)
  T2  best prior per gt, equal IoU      -> lowest prior index
  T3  several gts forcing one prior     -> highest gt index  (sequential index_put, last write wins)
  T4  hard-negative rank, equal CE      -> lower prior index (stable descending sort)
  T5  NMS candidate order, equal score  -> lower prior index (stable descending sort)
  T7  global top-k, equal score         -> earlier position in the class-major list
"""
from __future__ import annotations

from math import sqrt

import numpy as np
import torch
import torch.nn.functional as F

BG_CLASS = 20          # background is the LAST class (Losses.py:171, Util.py:26-27)
NUM_CLASSES = 21
POS_IOU = 0.5          # Losses.py:171
NEG_RATIO = 3          # Losses.py:189

SSD300 = dict(
    grids=[38, 19, 10, 5, 3, 1],
    scales=[0.1, 0.2, 0.375, 0.55, 0.725, 0.9],
    ratios=[[1., 2., 0.5], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333],
            [1., 2., 3., 0.5, .333], [1., 2., 0.5], [1., 2., 0.5]],
)
# SSD512-style stress table (BASELINE.json configs[4]): 24 564 priors.
SSD512 = dict(
    grids=[64, 32, 16, 8, 4, 2, 1],
    scales=[0.07, 0.15, 0.3, 0.45, 0.6, 0.75, 0.9],
    ratios=[[1., 2., 0.5], [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333],
            [1., 2., 3., 0.5, .333], [1., 2., 3., 0.5, .333], [1., 2., 0.5], [1., 2., 0.5]],
)


# --------------------------------------------------------------------------- priors
def make_priors(grids=None, scales=None, ratios=None) -> torch.Tensor:
    """Prior table [P,4] cx,cy,w,h float32 (Util.py:105-137).

    Order: level -> row (cy) -> column (cx) -> ratio, with one extra square prior
    of scale sqrt(s_k * s_{k+1}) directly after ratio 1 (1.0 on the last level).
    Math is Python float64, cast to float32, then clamped to [0,1] *in cxcywh*.
    """
    cfg = SSD300
    grids = cfg["grids"] if grids is None else grids
    scales = cfg["scales"] if scales is None else scales
    ratios = cfg["ratios"] if ratios is None else ratios
    rows = []
    for lvl, g in enumerate(grids):
        gf = float(g)
        nxt = sqrt(scales[lvl] * scales[lvl + 1]) if lvl + 1 < len(scales) else 1.
        for i in range(int(g)):
            for j in range(int(g)):
                cx = (j + 0.5) / gf
                cy = (i + 0.5) / gf
                for a in ratios[lvl]:
                    rows.append([cx, cy, scales[lvl] * sqrt(a), scales[lvl] / sqrt(a)])
                    if a == 1.:
                        rows.append([cx, cy, nxt, nxt])
    t = torch.tensor(rows, dtype=torch.float64).to(torch.float32)
    return t.clamp_(0, 1)


def cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    """Util.py:93-96: (c - wh/2, c + wh/2)."""
    half = b[:, 2:] / 2.
    return torch.cat((b[:, :2] - half, b[:, :2] + half), dim=1)


def xyxy_to_cxcywh(b: torch.Tensor) -> torch.Tensor:
    """Util.py:57-63 (numpy float32 round trip): ((x2+x1)/2, (y2+y1)/2, x2-x1, y2-y1)."""
    a = b.detach().cpu().numpy().astype(np.float32, copy=False)
    x1, y1, x2, y2 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    out = np.stack(((x2 + x1) / np.float32(2.), (y2 + y1) / np.float32(2.), x2 - x1, y2 - y1), axis=1)
    return torch.from_numpy(np.ascontiguousarray(out, dtype=np.float32))


def encode(cxcywh: torch.Tensor, pri: torch.Tensor) -> torch.Tensor:
    """Util.py:98-102: g_c = (c - pc) / (pwh / 10); g_wh = log(wh / pwh) * 5."""
    return torch.cat([(cxcywh[:, :2] - pri[:, :2]) / (pri[:, 2:] / 10),
                      torch.log(cxcywh[:, 2:] / pri[:, 2:]) * 5], 1)


def decode(g: torch.Tensor, pri: torch.Tensor) -> torch.Tensor:
    """Util.py:86-91: c = g_c * pwh / 10 + pc; wh = exp(g_wh / 5) * pwh."""
    return torch.cat([g[:, :2] * pri[:, 2:] / 10 + pri[:, :2],
                      torch.exp(g[:, 2:] / 5) * pri[:, 2:]], 1)


# --------------------------------------------------------------------------- IoU
def iou_matrix(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Dense jaccard [n1,n2] of xyxy boxes (Util.py:252-265 + 288-301).

    inter = prod(clamp(min(hi) - max(lo), 0)); union = (area_a + area_b) - inter.
    """
    lo = torch.max(a[:, None, :2], b[None, :, :2])
    hi = torch.min(a[:, None, 2:], b[None, :, 2:])
    d = torch.clamp(hi - lo, min=0)
    inter = d[..., 0] * d[..., 1]
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = area_a[:, None] + area_b[None, :] - inter
    return inter / union


# --------------------------------------------------------------------------- matching
def match_image(iou: torch.Tensor, gt_cls: torch.Tensor, thr: float = POS_IOU, bg: int = BG_CLASS):
    """Single-image match (Util.py:333-352; batched twin Losses.py:150-171).

    Returns (cls_per_prior float32 [P], obj_per_prior int64 [P] local gt index,
    overlap [P] after the forced override, best_prior_per_gt int64 [G]).
    """
    overlap, obj = iou.max(dim=0)                  # T1: first maximal gt
    _, best_prior = iou.max(dim=1)                 # T2: first maximal prior
    overlap = overlap.clone()
    obj = obj.clone()
    for g in range(iou.shape[0]):                  # T3: sequential, last write wins
        obj[best_prior[g]] = g
        overlap[best_prior[g]] = 1.
    cls = gt_cls.to(torch.float32)[obj]
    cls[overlap < thr] = bg                        # positive <=> not (iou < thr)
    return cls, obj, overlap, best_prior


def match_batch(gt_boxes, gt_cls, pri_xyxy, thr: float = POS_IOU, bg: int = BG_CLASS):
    """Batched match with GLOBAL gt indices (Losses.py:150-171).

    Returns dict: cls int64 [B,P], obj int64 [B,P] (index into the concatenated
    gt list), pos bool [B,P], npos int64 [B], best_prior int64 [sum G], off [B+1].
    """
    off = np.zeros(len(gt_boxes) + 1, dtype=np.int64)
    off[1:] = np.cumsum([int(b.shape[0]) for b in gt_boxes])
    cls_l, obj_l, bp_l = [], [], []
    for i, (bx, cl) in enumerate(zip(gt_boxes, gt_cls)):
        iou = iou_matrix(bx.view(-1, 4), pri_xyxy)
        c, o, _, bp = match_image(iou, cl, thr, bg)
        cls_l.append(c.to(torch.int64))
        obj_l.append(o + int(off[i]))
        bp_l.append(bp)
    cls = torch.stack(cls_l)
    pos = cls != bg
    return dict(cls=cls, obj=torch.stack(obj_l), pos=pos, npos=pos.sum(dim=1),
                best_prior=torch.cat(bp_l), off=off)


# --------------------------------------------------------------------------- loss
def cross_entropy_rows(conf: torch.Tensor, cls: torch.Tensor) -> torch.Tensor:
    """Per-prior CE [B,P] (Losses.py:184-185)."""
    b, p, c = conf.shape
    return F.cross_entropy(conf.reshape(-1, c), cls.reshape(-1), reduction='none').view(b, p)


def mine_hard_negatives(cce: torch.Tensor, pos: torch.Tensor, ratio: int = NEG_RATIO) -> torch.Tensor:
    """Mined-negative mask [B,P] under rule T4 (Losses.py:188-195).

    Positives take part in the ranking with value 0 (Losses.py:190); the k = ratio*npos_i
    highest entries of each row are taken, ties broken towards the lower prior index.
    Entries that are positives are removed from the returned mask (they carry no
    mined loss and no extra gradient).
    """
    v = cce.detach().clone()
    v[pos] = 0.
    order = torch.sort(v, dim=1, descending=True, stable=True).indices
    k = (ratio * pos.sum(dim=1)).clamp(max=v.shape[1])
    rank = torch.empty_like(order)
    ar = torch.arange(v.shape[1]).expand_as(order)
    rank.scatter_(1, order, ar)
    return (rank < k[:, None]) & ~pos


def multibox_loss(loc, conf, gt_boxes, gt_cls, pri_cxcywh, pri_xyxy=None,
                  thr: float = POS_IOU, ratio: int = NEG_RATIO, bg: int = BG_CLASS):
    """Batch multibox loss with stage taps (Losses.py:119-199).

    loc loss  = sum |loc_pos - enc| / (4 * Npos_total)      (nn.L1Loss mean, Losses.py:147,182)
    conf loss = (sum_pos CE + sum_mined CE) / Npos_total     (Losses.py:197)
    Returns dict with the two scalars (differentiable when loc/conf require grad)
    and the intermediate integer results.
    """
    if pri_xyxy is None:
        pri_xyxy = cxcywh_to_xyxy(pri_cxcywh)
    m = match_batch(gt_boxes, gt_cls, pri_xyxy, thr, bg)
    pos = m["pos"]
    gt_cxcywh = xyxy_to_cxcywh(torch.cat([b.view(-1, 4) for b in gt_boxes]))
    bs = loc.shape[0]
    tgt = encode(gt_cxcywh[m["obj"]][pos], pri_cxcywh.unsqueeze(0).expand(bs, -1, -1)[pos])
    loc_loss = torch.nn.functional.l1_loss(loc[pos], tgt)
    cce = cross_entropy_rows(conf, m["cls"])
    mined = mine_hard_negatives(cce, pos, ratio)
    npos_total = pos.sum().float()
    conf_loss = (cce[mined].sum().float() + cce[pos].sum().float()) / npos_total
    out = dict(m)
    out.update(loc_loss=loc_loss, conf_loss=conf_loss, cce=cce.detach(), mined=mined,
               target=tgt.detach(), npos_total=int(pos.sum()))
    return out


def multibox_grads(loc, conf, res, gout_loc: float = 1.0, gout_conf: float = 1.0):
    """Closed-form gradients of ``multibox_loss`` (what autograd gives through Losses.py:182-197):
    dloc = sign(loc - enc) / (4 Npos) on positives; dconf = (softmax - onehot) / Npos on
    positives and mined negatives; zero elsewhere."""
    pos, mined, cls = res["pos"], res["mined"], res["cls"]
    n = float(res["npos_total"])
    gl = torch.zeros_like(loc)
    gl[pos] = torch.sign(loc.detach()[pos] - res["target"]) * (gout_loc / (4.0 * n))
    sel = pos | mined
    gc = torch.zeros_like(conf)
    sm = torch.softmax(conf.detach()[sel], dim=1)
    sm[torch.arange(sm.shape[0]), cls[sel]] -= 1.0
    gc[sel] = sm * (gout_conf / n)
    return gl, gc


def ssd_reference_style(outputs, tr_classes, tr_bboxs, pri_cxcywh, pri_xyxy):
    """The reference's own op sequence for ``ssd()`` (Losses.py:119-199), kept as close to
    its cost profile as a restatement can be: one [sum G, P] IoU matrix, a Python loop of
    per-image dim-0 max, a second loop for the forced match, boolean-mask gathers and a FULL
    descending sort for the mining.  Used as the CPU baseline by bench.py and as the autograd
    check of ``multibox_grads``.  Returns (loc_loss, conf_loss)."""
    loc, conf = outputs
    bs, p = loc.shape[0], loc.shape[1]
    allb = torch.cat(tr_bboxs).view(-1, 4)
    iou = iou_matrix(allb, pri_xyxy)
    off = np.array([0] + [int(b.shape[0]) for b in tr_bboxs]).cumsum()
    ov_l, obj_l = [], []
    for i in range(bs):
        ov, ob = iou[off[i]:off[i + 1], :].max(dim=0)
        ov_l.append(ov)
        obj_l.append(ob + int(off[i]))
    _, best_prior = iou.max(dim=1)
    overlap, obj = torch.stack(ov_l), torch.stack(obj_l).long()
    gt_cxcywh, gt_c = xyxy_to_cxcywh(allb), torch.cat(tr_classes)
    for i in range(bs):
        sel = best_prior[off[i]:off[i + 1]]
        obj[i, :][sel] = torch.arange(int(off[i]), int(off[i + 1])).long()
        overlap[i, :][sel] = 1.
    cls = gt_c[obj]
    cls[overlap < POS_IOU] = BG_CLASS
    cls = cls.to(torch.int64)
    pos = cls != BG_CLASS
    tgt = encode(gt_cxcywh[obj][pos, :], pri_cxcywh.unsqueeze(0).repeat_interleave(bs, 0)[pos])
    loc_loss = torch.nn.L1Loss()(loc[pos, :], tgt)
    cce = F.cross_entropy(conf.view(-1, conf.shape[-1]), cls.view(-1), reduction='none').view(bs, p)
    pos_loss = cce[pos]
    c1 = cce.clone()
    c1[pos] = 0.
    c1, _ = c1.sort(dim=1, descending=True)
    hn = torch.arange(p).unsqueeze(0).expand_as(c1) < (NEG_RATIO * pos.sum(dim=1)).unsqueeze(1)
    conf_loss = (c1[hn].sum().float() + pos_loss.sum().float()) / pos.sum().float()
    return loc_loss, conf_loss


def ssd_per_image_mean(outputs, tr_classes, tr_bboxs, pri_cxcywh, pri_xyxy):
    """Legacy ``ssd_old`` / ``ssd1`` (Losses.py:100-117, 201-225): per-image losses with
    per-image normalisation, averaged over the batch.  Returns (loc_loss, conf_loss)."""
    loc, conf = outputs
    bs = len(tr_bboxs)
    lb = lc = 0.0
    for i in range(bs):
        r = multibox_loss(loc[i:i + 1], conf[i:i + 1], [tr_bboxs[i]], [tr_classes[i]], pri_cxcywh, pri_xyxy)
        lb = lb + r["loc_loss"]
        lc = lc + r["conf_loss"]
    return lb / bs, lc / bs


# --------------------------------------------------------------------------- detect
def nms_sorted(boxes_xyxy: torch.Tensor, thr: float) -> torch.Tensor:
    """Greedy NMS over boxes already sorted by descending score (Losses.py:41-55).
    Returns the keep mask.  A box suppresses every LATER-or-earlier box with
    IoU >= thr unless it is itself suppressed; a box never suppresses itself."""
    n = boxes_xyxy.shape[0]
    iou = iou_matrix(boxes_xyxy, boxes_xyxy)
    hit = (iou >= thr).numpy()
    sup = np.zeros(n, dtype=bool)
    for i in range(n):
        if sup[i]:
            continue
        sup |= hit[i]
        sup[i] = False
    return torch.from_numpy(~sup)


def detect_from_scores(boxes_cxcywh: torch.Tensor, probs: torch.Tensor, min_score: float, iou_thr: float,
                       top_k: int, num_fg: int = 20):
    """Stage-isolated detect: per-class threshold / sort / NMS / global top-k on GIVEN decoded
    boxes [P,4] (cxcywh) and probabilities [P,C] (Losses.py:27-81).

    Returns (boxes_xyxy [K,4], classes int64 [K], probs [K], prior_ids int64 [K]); K may be 0.
    Order: class-major (each class by descending score, T5) unless more than ``top_k``
    survive, then globally by descending score (stable, T7), truncated."""
    out_b, out_p, out_c, out_i = [], [], [], []
    for c in range(num_fg):
        pc = probs[:, c]
        cand = (pc >= min_score).nonzero().flatten()
        if cand.numel() == 0:
            continue
        sp, order = torch.sort(pc[cand], dim=0, descending=True, stable=True)
        ids = cand[order]
        bx = boxes_cxcywh[ids]
        keep = nms_sorted(cxcywh_to_xyxy(bx), iou_thr)
        out_b.append(bx[keep])
        out_p.append(sp[keep])
        out_c.append(torch.full((int(keep.sum()),), c, dtype=torch.int64))
        out_i.append(ids[keep])
    if not out_b:
        z = torch.zeros
        return z((0, 4)), z((0,), dtype=torch.int64), z((0,)), z((0,), dtype=torch.int64)
    b = cxcywh_to_xyxy(torch.cat(out_b))
    p, c, i = torch.cat(out_p), torch.cat(out_c), torch.cat(out_i)
    if b.shape[0] > top_k:
        p, order = torch.sort(p, dim=0, descending=True, stable=True)
        p, order = p[:top_k], order[:top_k]
        b, c, i = b[order], c[order], i[order]
    return b, c, p, i


def detect_image(loc: torch.Tensor, conf: torch.Tensor, pri_cxcywh: torch.Tensor, min_score: float = 0.2,
                 iou_thr: float = 0.45, top_k: int = 200, num_fg: int = 20):
    """Single-image ``inference`` without the drawing / pixel scaling (Losses.py:11-81):
    decode all priors, softmax over classes, then ``detect_from_scores``.  Boxes are
    fractional xyxy and are NOT clamped."""
    boxes = decode(loc, pri_cxcywh)
    probs = F.softmax(conf, dim=1)
    return detect_from_scores(boxes, probs, min_score, iou_thr, top_k, num_fg)


# --------------------------------------------------------------------------- evaluation
def voc_ap(det_boxes, det_classes, det_scores, gt_boxes, gt_classes, num_fg: int = 20, iou_thr: float = 0.5):
    """11-point interpolated VOC average precision per class (``get_map``, Util.py:783-885).

    Inputs are per-image lists of tensors.  Per class: detections ranked by descending score (stable: ties -> lower
    detection index, rule M1; the reference's sort leaves ties open); in that order a detection is a true positive
    iff the gt of the same image and class with the highest IoU (first one on ties, M2) has IoU > iou_thr and has
    not been claimed yet (Util.py:855-868).  Precision in float64; recall = cumTP * float32(1 / #gt) (see below);
    recall levels = the float32 values of ``torch.arange(0, 1.1, 0.1)`` (Util.py:875).  Returns float64 [num_fg]."""
    det_img = torch.cat([torch.full((len(b),), i, dtype=torch.int64) for i, b in enumerate(det_boxes)])
    db = torch.cat([b.reshape(-1, 4) for b in det_boxes]).float()
    dc = torch.cat([c.reshape(-1) for c in det_classes]).long()
    ds = torch.cat([s.reshape(-1) for s in det_scores]).float()
    gt_img = torch.cat([torch.full((len(b),), i, dtype=torch.int64) for i, b in enumerate(gt_boxes)])
    gb = torch.cat([b.reshape(-1, 4) for b in gt_boxes]).float()
    gc = torch.cat([c.reshape(-1) for c in gt_classes]).long()
    claimed = np.zeros(gb.shape[0], dtype=bool)
    levels = torch.arange(0, 1.1, 0.1).double().numpy()
    ap = np.zeros(num_fg, dtype=np.float64)
    for c in range(num_fg):
        idx = (dc == c).nonzero().flatten()
        if idx.numel() == 0:
            continue
        order = torch.sort(ds[idx], descending=True, stable=True).indices
        idx = idx[order]
        nobj = int((gc == c).sum())
        tp = np.zeros(idx.numel(), dtype=np.float64)
        for r, d in enumerate(idx.tolist()):
            g = ((gt_img == det_img[d]) & (gc == c)).nonzero().flatten()
            if g.numel() == 0:
                continue
            iou = iou_matrix(db[d:d + 1], gb[g])[0]
            ov, j = iou.max(dim=0)
            gi = int(g[j])
            if float(ov) > iou_thr and not claimed[gi]:
                claimed[gi] = True
                tp[r] = 1.0
        ctp = tp.cumsum()
        precision = ctp / np.arange(1, len(tp) + 1, dtype=np.float64)
        # Util.py:872 divides a numpy array by a 0-dim int64 TENSOR: torch evaluates that as reciprocal(tensor) * array,
        # and the reciprocal of an integer tensor is float32 - so recall = cumTP * float32(1/#gt), a hair above k/#gt.
        with np.errstate(divide="ignore", invalid="ignore"):
            recall = ctp * np.float64(np.float32(1.0) / np.float32(nobj))
        vals = [precision[recall >= lv].max() if (recall >= lv).any() else 0.0 for lv in levels]
        ap[c] = float(np.mean(vals))
    return ap


# --------------------------------------------------------------------------- gt collate
def collate_gt(boxes, classes, difficult=None, keep_difficult=True, img_wh=None):
    """The reference's gt handling between the dataset and the loss, restated with its own torch ops:
    ``Dataset.py:28-30`` (drop difficult boxes), ``Dataset.py:35-36`` (``bboxes / FloatTensor([w, h, w, h])``),
    ``Losses.py:129-130`` (``torch.cat`` + cumulative offsets).  Returns (boxes [sumG,4], classes [sumG] float32,
    offsets int32 [B+1]); an image left without a box raises IndexError (``Losses.py:153`` fails on it)."""
    out_b, out_c, off = [], [], [0]
    for i, (b, c) in enumerate(zip(boxes, classes)):
        b = torch.as_tensor(b, dtype=torch.float32).reshape(-1, 4)
        c = torch.as_tensor(c, dtype=torch.float32).reshape(-1)
        if difficult is not None and not keep_difficult:
            keep = torch.as_tensor(difficult[i]).reshape(-1) == 0
            b, c = b[keep], c[keep]
        if img_wh is not None:
            w, h = float(img_wh[i][0]), float(img_wh[i][1])
            b = b / torch.FloatTensor([w, h, w, h]).unsqueeze(0)
        if b.shape[0] == 0:
            raise IndexError(f"image {i} has no ground-truth box")
        out_b.append(b)
        out_c.append(c)
        off.append(off[-1] + b.shape[0])
    return torch.cat(out_b), torch.cat(out_c), torch.tensor(off, dtype=torch.int32)

