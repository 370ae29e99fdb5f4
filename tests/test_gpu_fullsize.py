"""GPU, BASELINE.json's full sizes (where the Python oracle would take minutes): size-independent properties.

  configs[3]  SSD300 head, batch 256            match + loss fwd/bwd
  configs[2]  SSD300 detect, batch 64, bias +6  decode + NMS + top-200
  configs[4]  24 564 priors, 100 gt/image, batch 128   match + loss, decode + NMS (the configuration's stated size)
Every image of every batch is covered by the properties, AND every image is compared with the oracle: the whole
256-image (300-image, 128-image stress) batch of the training head - class map, CE, mined sets with a rank-boundary
proof, losses, every gradient row - and all 64 / 256 images of the detect batches (32 of the 128 stress images: the
oracle's per-class NMS loop needs ~0.5 s per 24 564-prior image), differing detections only with a boundary proof.
"""
import pytest
import torch

from oracle import ssd_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _head(pri):
    from objectdetection_ssd_b200.head import MultiboxHead
    return MultiboxHead(pri, "cuda")


def _loss_properties(pri, loc, conf, tb, tc, check_images):
    from objectdetection_ssd_b200.head import PackedGT
    head = _head(pri)
    P = pri.shape[0]
    B = loc.shape[0]
    gt = PackedGT(tb, tc, head.dev)
    l, c = loc.cuda(), conf.cuda()
    out = head.loss(l, c, gt, with_grads=True, taps=True)
    torch.cuda.synchronize()
    npos = out["npos"][:B].long()
    cls = out["cls_u8"].long()
    pos = cls != 20
    mined = H.unpack_mask(out["mined_mask"], P).cuda()
    ce = out["ce"]
    # positives: count, every gt owns at least one (forced match, Losses.py:164-167), classes come from the image's gts
    assert torch.equal(pos.sum(1), npos) and int(out["npos"][B]) == int(npos.sum())
    off = gt.off_host
    for b in check_images:
        bp = out["best_prior"][off[b]:off[b + 1]].long()
        assert pos[b, bp].all(), "a forced prior must be positive"
        assert set(cls[b][pos[b]].tolist()) <= set(int(x) for x in tc[b].tolist())
    # mining: exactly min(3*npos, #negatives) per image, disjoint from positives, and it IS the top-k of the background CE
    assert not (mined & pos).any()
    k = torch.minimum(3 * npos, (~pos).sum(1))
    assert torch.equal(mined.sum(1), k)
    neg_ce = ce.masked_fill(pos, 0.0)
    worst_in = torch.where(mined, neg_ce, torch.full_like(neg_ce, float("inf"))).min(1).values
    best_out = torch.where(~mined & ~pos, neg_ce, torch.full_like(neg_ce, -1.0)).max(1).values
    assert (worst_in >= best_out).all(), "a non-mined negative has a larger CE than a mined one"
    # loss == recomputation from the taps (Losses.py:197), fp64
    n = float(out["npos"][B])
    conf_loss = (ce.double()[pos].sum() + ce.double()[mined].sum()) / n
    assert abs(out["losses"][1].item() - conf_loss.item()) <= 1e-5 * conf_loss.item()
    assert abs(out["sums"][1].item() - conf_loss.item() * n) <= 1e-6 * conf_loss.item() * n
    # gradients: non-zero rows == positives + mined; each such conf row sums to 0 (softmax - onehot); loc rows on positives only
    gc, gl = out["grad_conf"], out["grad_loc"]
    sel = pos | mined
    assert torch.equal(gc.abs().sum(-1) != 0, sel)
    assert gc.sum(-1).abs().max().item() <= 1e-6 / max(n, 1.0) * 10
    assert torch.equal(gl.abs().sum(-1) != 0, pos) or (gl.abs().sum(-1) != 0).sum() <= pos.sum()
    assert (gl.abs()[pos] <= 1.0 / (4 * n) * (1 + 1e-6)).all()
    # the true class has a negative gradient, every other class a non-negative one
    rows = gc[sel]
    tgt = cls[sel]
    assert (rows.gather(1, tgt[:, None]) <= 0).all()
    assert (rows.scatter(1, tgt[:, None], 0.0) >= 0).all()
    # a slice of the batch against the oracle (class map exact, CE to 1e-5)
    for b in check_images:
        ref = O.multibox_loss(loc[b:b + 1], conf[b:b + 1], [tb[b]], [tc[b]], pri)
        assert torch.equal(cls[b].cpu(), ref["cls"][0])
        assert torch.allclose(ce[b].cpu(), ref["cce"][0], rtol=1e-5, atol=1e-6)
        assert int((mined[b].cpu() ^ ref["mined"][0]).sum()) <= 2
    # image order does not matter (per-image independence: what makes the batch shardable)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    out2 = head.loss(l[perm.cuda()], c[perm.cuda()], PackedGT([tb[i] for i in perm], [tc[i] for i in perm], head.dev),
                     with_grads=True)
    torch.cuda.synchronize()
    assert torch.equal(out2["cls_u8"], out["cls_u8"][perm.cuda()])
    assert torch.equal(out2["grad_conf"], out["grad_conf"][perm.cuda()])
    assert abs(out2["losses"][1].item() - out["losses"][1].item()) <= 1e-6 * out["losses"][1].item()


def _whole_batch_vs_oracle(pri, loc, conf, tb, tc):
    """every image of the batch against the oracle (the same checks the small-batch loss tests make)"""
    from tests.test_gpu_loss import _check_loss
    out, ref = _check_loss(pri, loc, conf, tb, tc)
    assert torch.equal(out["cls_u8"].long().cpu(), ref["cls"]), "class map of every image"
    assert torch.equal(out["npos"][:loc.shape[0]].long().cpu(), ref["npos"])


def test_train_head_batch256_properties():
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(71, 256, pri.shape[0])
    _loss_properties(pri, loc, conf, tb, tc, check_images=[0, 100, 255])
    _whole_batch_vs_oracle(pri, loc, conf, tb, tc)


def test_train_head_batch300_exceeds_coresident_grid():
    """B = 300 > 2 CTAs x 148 SMs: the cooperative two-kernel step does not fit and ssdhead_multibox_step falls back
    to the three-kernel route (stream, finaliser, mining) by itself; same properties must hold."""
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(75, 300, pri.shape[0])
    _loss_properties(pri, loc, conf, tb, tc, check_images=[0, 299])
    _whole_batch_vs_oracle(pri, loc, conf, tb, tc)


def test_train_head_stress_24564_priors_100_gt():
    pri = H.priors("ssd512")
    loc, conf, tb, tc = H.train_inputs(72, 128, pri.shape[0], min_gt=100, max_gt=100)
    _loss_properties(pri, loc, conf, tb, tc, check_images=[0, 127])
    _whole_batch_vs_oracle(pri, loc, conf, tb, tc)


def _detect_properties(pri, loc, conf, min_score, top_k, check_images, oracle_images=None):
    from objectdetection_ssd_b200.head import detect, detect_from_scores
    head = _head(pri)
    B, P = loc.shape[0], pri.shape[0]
    out = detect(head, loc, conf, min_score, 0.45, top_k)
    torch.cuda.synchronize()
    n = H.compare_detect_with_oracle(out, loc, conf, pri, min_score, 0.45, top_k, range(B) if oracle_images is None else oracle_images)
    print(f"detect B={B}: {n} detections differ from the oracle's, all with a boundary proof")
    cnt = out["cnt"].cpu()
    assert (cnt >= 0).all() and (cnt <= top_k).all()
    pxy = None
    for b in range(B):
        k = int(cnt[b])
        prob, cls, prior, boxes = out["prob"][b, :k], out["cls"][b, :k], out["prior"][b, :k].long(), out["boxes"][b, :k]
        assert (prob >= min_score).all() and ((cls >= 0) & (cls < 20)).all() and ((prior >= 0) & (prior < P)).all()
        if k == top_k:
            assert (prob[:-1] >= prob[1:]).all(), "a full list is sorted by descending score (Losses.py:77-81)"
        key = cls.long() * P + prior
        assert key.unique().numel() == k, "a (class, prior) pair appears once"
        if b in check_images:
            # no two kept boxes of one class overlap by >= thr (NMS invariant), checked with the oracle's exact IoU
            bc, cc = boxes.cpu(), cls.cpu()
            for c in cc.unique().tolist():
                sel = bc[cc == c]
                iou = O.iou_matrix(sel, sel)
                iou.fill_diagonal_(0)
                assert (iou < 0.45).all()
            # decoded boxes agree with the oracle's decode of those priors
            ref_boxes = O.cxcywh_to_xyxy(O.decode(loc[b], pri))[prior.cpu()]
            assert torch.allclose(bc, ref_boxes, rtol=1e-5, atol=1e-6)
    # idempotence: detections of image b, fed back as the only candidates, all survive (stage-isolated entry)
    b = check_images[0]
    k = int(cnt[b])
    probs = torch.zeros(1, P, 21)
    probs[0, :, 20] = 1.0
    probs[0, out["prior"][b, :k].long().cpu(), out["cls"][b, :k].long().cpu()] = out["prob"][b, :k].cpu()
    boxes_cx = O.decode(loc[b], pri).unsqueeze(0)
    again = detect_from_scores(head, boxes_cx, probs, min_score, 0.45, top_k)
    torch.cuda.synchronize()
    assert int(again["cnt"][0]) == k
    a = set(zip(again["cls"][0, :k].tolist(), again["prior"][0, :k].tolist()))
    assert a == set(zip(out["cls"][b, :k].tolist(), out["prior"][b, :k].tolist()))


def test_detect_batch64_bias6_properties():
    pri = H.priors()
    loc, conf = H.detect_inputs(73, 64, pri.shape[0], bg_bias=6.0)
    _detect_properties(pri, loc, conf, 0.01, 200, check_images=[0, 63])


def test_detect_batch256_bias6_properties():
    """north_star's target size for decode + NMS."""
    pri = H.priors()
    loc, conf = H.detect_inputs(76, 256, pri.shape[0], bg_bias=6.0)
    _detect_properties(pri, loc, conf, 0.01, 200, check_images=[0, 255])


def test_detect_stress_24564_priors():
    pri = H.priors("ssd512")
    loc, conf = H.detect_inputs(74, 128, pri.shape[0], bg_bias=6.0)
    _detect_properties(pri, loc, conf, 0.01, 200, check_images=[0], oracle_images=range(0, 128, 4))
