"""Shared input builders for the parity tests (seeded, identical bits for oracle and kernels)."""
import numpy as np
import torch

from objectdetection_ssd_b200 import synth
from oracle import ssd_oracle as O


def priors(kind="ssd300"):
    if kind == "ssd300":
        return O.make_priors()
    if kind == "ssd512":
        return O.make_priors(**O.SSD512)
    raise ValueError(kind)


def train_inputs(seed, B, P, min_gt=1, max_gt=10, C=21):
    gb, gc = synth.make_gt(seed, B, min_gt, max_gt)
    loc, conf = synth.make_head(seed, B, P, C)
    tb = [torch.from_numpy(b) for b in gb]
    tc = [torch.from_numpy(c) for c in gc]
    return torch.from_numpy(loc), torch.from_numpy(conf), tb, tc


def detect_inputs(seed, B, P, bg_bias=6.0, C=21):
    loc, conf = synth.make_head(seed, B, P, C, loc_scale=0.5, bg_bias=bg_bias)
    return torch.from_numpy(loc), torch.from_numpy(conf)


def unpack_mask(mask_i32: torch.Tensor, P: int) -> torch.Tensor:
    """uint32 bit mask [B, ceil(P/32)] (stored as int32) -> bool [B,P]."""
    m = mask_i32.cpu().numpy().view(np.uint32)
    bits = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(bool)
    return torch.from_numpy(bits.reshape(m.shape[0], -1)[:, :P])


# ---------------------------------------------------------------------------------------------------------------
# Boundary proof for the FUSED detect path (own softmax / decode on the device).  CUDA's ex2.approx / expf and ATen's
# exp differ in the last bits, so a detection may differ from the oracle's - but only for a reason:
#   (a) its OWN probability is within PROB_ULP ulp of `min_score` (the candidate itself flips; it ranks last in its
#       class, so nothing else can change because of it);
#   (b) its class holds a pair (box kept by either side, candidate) whose IoU is within IOU_REL of the NMS threshold (a
#       suppression flips, and the greedy sweep of THAT class may cascade);
#   (c) more than top_k survive and it sits at the global cut: its probability is within PROB_ULP ulp of the k-th
#       largest, or it is at / below the cut while some class has a (b) event (one more or one fewer box kept above
#       moves the cut).
# A difference without such an event is a genuine error.  Stated tolerances: the streaming softmax uses
# ex2.approx.ftz, relative error <= (2 + 1.16 |x - max|) ulp per term (csrc/common.cuh), i.e. <= 16 ulp on a
# probability at |x - max| <= 10; decoded boxes differ by a few ulp (expf), their IoU by < 1e-5 relative.
PROB_ULP = 16
IOU_REL = 1e-5


def _ulps(a: float, b: float) -> float:
    import math
    m = max(abs(a), abs(b), 1e-30)
    return abs(a - b) / (2.0 ** (math.floor(math.log2(m)) - 23))


def explain_detect_mismatches(loc_i, conf_i, pri, min_score, iou_thr, top_k, ours, ref):
    """``ours`` / ``ref``: sets of (class, prior) detections of one image.  Returns (n_mismatches, unexplained list)."""
    import torch.nn.functional as F
    diff = (ours - ref) | (ref - ours)
    if not diff:
        return 0, []
    probs = F.softmax(conf_i, dim=1)
    boxes = O.cxcywh_to_xyxy(O.decode(loc_i, pri))
    iou_fragile = set()
    for c in {c for c, _ in diff}:
        # only a box that one of the two sides KEPT can suppress anything: pairs (kept box, any candidate of the class)
        cand = torch.nonzero(probs[:, c] >= min_score * (1 - 1e-5)).flatten()
        kept = torch.tensor(sorted({p for cc, p in (ours | ref) if cc == c}), dtype=torch.long)
        if cand.numel() > 0 and kept.numel() > 0:
            iou = O.iou_matrix(boxes[kept], boxes[cand])
            iou[kept[:, None] == cand[None, :]] = 0
            if bool(((iou - iou_thr).abs() <= IOU_REL * iou_thr).any()):
                iou_fragile.add(c)
    cut = None
    if len(ref) >= top_k or len(ours) >= top_k:
        ps = sorted((float(probs[p, c]) for c, p in ref), reverse=True)
        cut = ps[min(len(ps), top_k) - 1]                       # k-th largest probability the oracle reported
    unexplained = []
    for c, p in diff:
        pr = float(probs[p, c])
        own = _ulps(pr, min_score) <= PROB_ULP
        at_cut = cut is not None and (_ulps(pr, cut) <= PROB_ULP or (iou_fragile and pr <= cut * (1 + 1e-5)))
        if not (own or c in iou_fragile or at_cut):
            unexplained.append((c, p))
    return len(diff), unexplained


def compare_detect_with_oracle(out, loc, conf, pri, min_score, iou_thr, top_k, images):
    """The device path's detections (own softmax + decode) of ``images`` against the oracle's: common detections agree to
    1e-5 in score and box, every differing detection needs a boundary proof (explain_detect_mismatches).  Returns the
    number of (explained) differences."""
    total = 0
    for i in images:
        rb, rc, rp, ri = O.detect_image(loc[i], conf[i], pri, min_score, iou_thr, top_k)
        k = int(out["cnt"][i])
        gp, gi = out["prob"][i, :k].cpu(), out["prior"][i, :k].cpu().long()
        gc, gb = out["cls"][i, :k].cpu().long(), out["boxes"][i, :k].cpu()
        ref = {(int(b), int(a)): j for j, (a, b) in enumerate(zip(ri, rc))}
        ours = {(int(b), int(a)): j for j, (a, b) in enumerate(zip(gi, gc))}
        common = sorted(set(ref) & set(ours))
        j = torch.tensor([ours[c] for c in common], dtype=torch.long)
        h = torch.tensor([ref[c] for c in common], dtype=torch.long)
        assert torch.allclose(gp[j], rp[h], rtol=1e-5, atol=1e-8), f"image {i}: scores"
        assert torch.allclose(gb[j], rb[h], rtol=1e-5, atol=1e-6), f"image {i}: boxes"
        n, unexplained = explain_detect_mismatches(loc[i], conf[i], pri, min_score, iou_thr, top_k, set(ours), set(ref))
        assert not unexplained, f"image {i}: {len(unexplained)} of {n} differing detections have no boundary proof: {unexplained[:5]}"
        total += n
    return total
