"""GPU: the drop-in call surface (Losses.py / Util.py names of the reference) against the oracle."""
import pytest
import torch

from oracle import ssd_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_util_box_functions():
    from objectdetection_ssd_b200 import Util
    pri = H.priors()
    g = torch.randn(8732, 4, generator=torch.Generator().manual_seed(1)) * 0.5
    dec = Util.gcxgcy_to_cxcy(g, pri)
    assert dec.is_cuda and torch.allclose(dec.cpu(), O.decode(g, pri), rtol=1e-6, atol=1e-7)
    xy = Util.xywh_to_xyxy(pri)
    assert torch.equal(xy.cpu(), O.cxcywh_to_xyxy(pri))
    back = Util.xyxy_to_xywh(xy)
    assert not back.is_cuda and torch.equal(back, O.xyxy_to_cxcywh(O.cxcywh_to_xyxy(pri)))
    boxes = O.cxcywh_to_xyxy(pri[100:400])
    enc = Util.get_offsets_coords(O.xyxy_to_cxcywh(boxes), pri[:300])
    assert torch.allclose(enc.cpu(), O.encode(O.xyxy_to_cxcywh(boxes), pri[:300]), rtol=1e-5, atol=1e-6)


def test_util_iou_and_match_from_matrix():
    from objectdetection_ssd_b200 import Util
    pri = H.priors()
    pxy = O.cxcywh_to_xyxy(pri)
    _, _, tb, tc = H.train_inputs(41, 3, pri.shape[0])
    for b, c in zip(tb, tc):
        j = Util.get_jaccard_tensor1(b, pxy)
        ref = O.iou_matrix(b, pxy)
        assert torch.equal(j.cpu(), ref)                                   # bit-exact IoU (T8)
        assert torch.equal(Util.get_jaccard_tensor11(b, pxy), ref)         # CPU twin
        assert torch.equal(Util.find_intersection(b.cuda(), pxy.cuda()).cpu(), Util.find_intersection(b, pxy))
        cls, obj = Util.map_prior_to_bb(j, c)
        rc, ro, _, _ = O.match_image(ref, c)
        assert torch.equal(cls.cpu(), rc) and torch.equal(obj.cpu(), ro)
    # T3 through the matrix entry: identical gts
    box = torch.tensor([[0.2, 0.2, 0.6, 0.7]]).repeat(3, 1)
    c = torch.tensor([3., 7., 5.])
    cls, obj = Util.map_prior_to_bb(O.iou_matrix(box, pxy), c)
    rc, ro, _, _ = O.match_image(O.iou_matrix(box, pxy), c)
    assert torch.equal(cls.cpu(), rc) and torch.equal(obj.cpu(), ro)


def test_losses_ssd_surface_and_variants():
    from objectdetection_ssd_b200 import Losses
    pri = H.priors()
    assert torch.equal(Losses.ancs_xywh, pri) and torch.equal(Losses.ancs_xyxy, O.cxcywh_to_xyxy(pri))
    loc, conf, tb, tc = H.train_inputs(42, 6, pri.shape[0])
    l = loc.cuda().requires_grad_(True)
    c = conf.cuda().requires_grad_(True)
    l1, l2 = Losses.ssd((l, c), [x.cuda() for x in tc], [x.cuda() for x in tb])
    assert l1.dim() == 0 and l2.dim() == 0 and l1.grad_fn is not None
    (l1 + l2).backward()
    ref = O.multibox_loss(loc, conf, tb, tc, pri)
    assert abs(l1.item() - ref["loc_loss"].item()) <= 1e-5 * ref["loc_loss"].item()
    assert abs(l2.item() - ref["conf_loss"].item()) <= 1e-5 * ref["conf_loss"].item()
    gl, gc = O.multibox_grads(loc, conf, ref)
    assert torch.allclose(l.grad.cpu(), gl, rtol=1e-5, atol=1e-9)
    assert torch.allclose(c.grad.cpu(), gc, rtol=1e-4, atol=1e-8)
    assert torch.equal(Losses.obj_forEach_prior___.cpu().long(), ref["cls"])     # debug tap (Losses.py:172-173)
    # inner function returns the pair swapped (Losses.py:199)
    a, b = Losses.ssd1_(loc.cuda(), conf.cuda(), tb, tc, None, None)
    assert abs(a.item() - l2.item()) < 1e-7 and abs(b.item() - l1.item()) < 1e-7
    # legacy per-image mean (Losses.py:100-117)
    o1, o2 = Losses.ssd_old((loc.cuda(), conf.cuda()), tc, tb)
    r1, r2 = O.ssd_per_image_mean((loc, conf), tc, tb, pri, O.cxcywh_to_xyxy(pri))
    assert abs(o1.item() - r1.item()) <= 1e-5 * r1.item() and abs(o2.item() - r2.item()) <= 1e-5 * r2.item()
    # no grad requested -> forward only
    with torch.no_grad():
        f1, f2 = Losses.ssd((loc.cuda(), conf.cuda()), tc, tb)
    assert abs(f1.item() - l1.item()) < 1e-7 and not f1.requires_grad


def test_losses_inference_surface(monkeypatch):
    from objectdetection_ssd_b200 import Losses
    pri = H.priors()
    loc, conf = H.detect_inputs(43, 2, pri.shape[0], bg_bias=8.0)
    monkeypatch.setattr(Losses, "_image_size", lambda phase, index: (500, 375))
    for i in range(2):
        boxes, cls, prob = Losses.inference(loc[i].cuda(), conf[i].cuda(), 0, toDraw=False, min_score=0.01)
        rb, rc, rp, ri = O.detect_image(loc[i], conf[i], pri, 0.01, 0.45, 200)
        assert boxes.shape[1] == 4 and cls.dtype == torch.int64
        scale = torch.tensor([500., 375., 500., 375.])
        # pair detections by (class, probability): every record of ours has its counterpart in the oracle's output with
        # the same box, or sits at a boundary (explained below through the prior ids of the batched front end)
        hit = (cls.cpu()[:, None] == rc[None, :]) & ((prob.cpu()[:, None] - rp[None, :]).abs() <= 1e-5 * rp[None, :])
        j = hit.float().argmax(1)
        ok = hit.any(1)
        assert torch.allclose(boxes.cpu()[ok], (rb * scale)[j[ok]], rtol=1e-5, atol=1e-3)
        det = Losses.inference_batch(loc[i:i + 1].cuda(), conf[i:i + 1].cuda(), top_k=200, min_score=0.01)
        k = int(det["cnt"][0])
        ours = {(int(c), int(p)) for c, p in zip(det["cls"][0, :k].cpu(), det["prior"][0, :k].cpu())}
        ref = {(int(c), int(p)) for c, p in zip(rc, ri)}
        assert k == boxes.shape[0] and int((~ok).sum()) <= len(ours ^ ref)
        n, unexplained = H.explain_detect_mismatches(loc[i], conf[i], pri, 0.01, 0.45, 200, ours, ref)
        assert not unexplained, f"image {i}: {unexplained[:5]} of {n} differing detections have no boundary proof"
    # nothing above the threshold -> three empty lists (Losses.py:62-63)
    conf0 = torch.zeros(pri.shape[0], 21)
    conf0[:, 20] = 30.0
    assert Losses.inference(loc[0].cuda(), conf0.cuda(), 0, toDraw=False) == ([], [], [])


def test_stress_prior_table_by_overwriting_module_global():
    """The 24 564-prior configuration is driven the way the reference would be: by replacing Losses.ancs_xywh."""
    from objectdetection_ssd_b200 import Losses
    pri = H.priors("ssd512")
    old = Losses.ancs_xywh
    try:
        Losses.ancs_xywh = pri
        loc, conf, tb, tc = H.train_inputs(44, 2, pri.shape[0], min_gt=100, max_gt=100)
        l1, l2 = Losses.ssd((loc.cuda(), conf.cuda()), tc, tb)
        ref = O.multibox_loss(loc, conf, tb, tc, pri)
        assert abs(l1.item() - ref["loc_loss"].item()) <= 1e-5 * ref["loc_loss"].item()
        assert abs(l2.item() - ref["conf_loss"].item()) <= 1e-5 * ref["conf_loss"].item()
    finally:
        Losses.ancs_xywh = old


def test_pinned_collate_arrays_own_their_block_and_staging_pool_is_recycled():
    """(high) arrays returned by collate_gt keep their page-locked block alive after the tuple is gone;
    (medium) ssd() with host lists recycles staging blocks instead of cudaHostAlloc / cudaFreeHost per call, and a
    block is not rewritten before the copy that read it has run."""
    import gc
    import numpy as np
    from objectdetection_ssd_b200 import head as HD
    from objectdetection_ssd_b200.collate import collate_gt
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    pri = H.priors()
    _, _, tb, tc = H.train_inputs(61, 9, pri.shape[0])
    gb, gcl, go = collate_gt(tb, tc)
    want = (np.concatenate([b.numpy() for b in tb]), np.concatenate([c.numpy() for c in tc]))
    gc.collect()
    junk = [collate_gt(tb, tc) for _ in range(16)]
    for j in junk:
        j[0][...] = -7.0
    del junk
    gc.collect()
    assert np.array_equal(gb, want[0]) and np.array_equal(gcl, want[1]) and go[-1] == want[0].shape[0]
    # staging pool: many ragged batches back to back on one stream; every packed gt must hold its own batch
    head = MultiboxHead(pri, "cuda")
    packed, wants = [], []
    for s in range(40):
        _, _, b, c = H.train_inputs(100 + s, 5 + s % 4, pri.shape[0])
        packed.append(PackedGT(b, c, head.dev))
        wants.append((torch.cat(b), torch.cat(c)))
    torch.cuda.synchronize()
    for g, (wb, wc) in zip(packed, wants):
        assert torch.equal(g.boxes.cpu(), wb) and torch.equal(g.classes.cpu(), wc)
        assert g.off.cpu().tolist() == g.off_host
    assert HD._staging_pool is not None and len(HD._staging_pool.free) <= 32


def test_util_box_functions_are_differentiable_like_the_reference():
    """Util.py:86-102 are plain torch expressions in the reference, so autograd flows through them."""
    from objectdetection_ssd_b200 import Util
    pri = H.priors()[:300]
    g = (torch.randn(300, 4, generator=torch.Generator().manual_seed(3)) * 0.5)
    w = torch.randn(300, 4, generator=torch.Generator().manual_seed(4))
    a = g.clone().requires_grad_(True)
    (Util.gcxgcy_to_cxcy(a, pri).cpu() * w).sum().backward()
    r = g.clone().requires_grad_(True)
    (O.decode(r, pri) * w).sum().backward()
    assert torch.allclose(a.grad, r.grad, rtol=1e-5, atol=1e-7)
    boxes = O.xyxy_to_cxcywh(O.cxcywh_to_xyxy(H.priors()[100:400])).clone()
    a = boxes.clone().requires_grad_(True)
    (Util.get_offsets_coords(a, pri).cpu() * w).sum().backward()
    r = boxes.clone().requires_grad_(True)
    (O.encode(r, pri) * w).sum().backward()
    assert torch.allclose(a.grad, r.grad, rtol=1e-5, atol=1e-6)
    a = boxes.clone().requires_grad_(True)
    (Util.xywh_to_xyxy(a).cpu() * w).sum().backward()
    r = boxes.clone().requires_grad_(True)
    (O.cxcywh_to_xyxy(r) * w).sum().backward()
    assert torch.allclose(a.grad, r.grad, rtol=1e-6, atol=1e-7)
    with torch.no_grad():
        assert not Util.gcxgcy_to_cxcy(a, pri).requires_grad


def test_inference_standalone_image_size_lookup(tmp_path):
    """inference(toDraw=False) without the reference's dataset lists: `index` as the image path (local get_img_sz,
    Util.py:226-228) or as a (w, h) pair - boxes scaled by (w, h, w, h) as Losses.py:87-89 does."""
    from PIL import Image
    from objectdetection_ssd_b200 import Losses, Util
    path = tmp_path / "img.png"
    Image.new("RGB", (500, 375)).save(path)
    assert tuple(Util.get_img_sz(str(path))) == (500, 375)
    pri = H.priors()
    loc, conf = H.detect_inputs(47, 1, pri.shape[0], bg_bias=8.0)
    b1, c1, p1 = Losses.inference(loc[0].cuda(), conf[0].cuda(), str(path), toDraw=False, min_score=0.05)
    b2, c2, p2 = Losses.inference(loc[0].cuda(), conf[0].cuda(), (500, 375), toDraw=False, min_score=0.05)
    b0, c0, p0 = Losses.inference(loc[0].cuda(), conf[0].cuda(), (1, 1), toDraw=False, min_score=0.05)
    assert b1.shape[0] > 0 and torch.equal(b1, b2) and torch.equal(c1, c2) and torch.equal(p1, p0)
    assert torch.equal(b1.cpu(), b0.cpu() * torch.tensor([500., 375., 500., 375.]))
