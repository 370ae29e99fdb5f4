"""GPU parity: match + multibox loss (fwd and gradients) against the CPU oracle.

Integer results (best prior per gt, object/class maps, positive counts, mined-negative sets)
must be bit-exact under tie rules T1-T4; losses, CE values and gradients agree to 1e-5 relative
(fp32; tolerance from BASELINE.json north_star).  Everything goes through the C ABI.
"""
import numpy as np
import pytest
import torch

from oracle import ssd_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

REL = 1e-5


def _head(pri):
    from objectdetection_ssd_b200.head import MultiboxHead
    return MultiboxHead(pri, "cuda")


def _packed(tb, tc, head):
    from objectdetection_ssd_b200.head import PackedGT
    return PackedGT(tb, tc, head.dev)


def _check_match(head, pri, tb, tc):
    gt = _packed(tb, tc, head)
    m = head.match(gt, want_maps=True)
    torch.cuda.synchronize()
    ref = O.match_batch(tb, tc, O.cxcywh_to_xyxy(pri))
    assert torch.equal(m["best_prior"][:gt.sumG].cpu().long(), ref["best_prior"]), "best prior per gt (T2)"
    assert torch.equal(m["cls"].cpu().long(), ref["cls"]), "class map"
    assert torch.equal(m["obj"].cpu().long(), ref["obj"]), "object map (T1/T3)"
    npos = m["npos"].cpu().long()
    assert torch.equal(npos[:-1], ref["npos"]), "positives per image"
    assert int(npos[-1]) == int(ref["npos"].sum())
    return gt, m, ref


@pytest.mark.parametrize("seed,B", [(1, 8), (2, 32)])
def test_match_ssd300_exact(seed, B):
    pri = H.priors()
    _, _, tb, tc = H.train_inputs(seed, B, pri.shape[0])
    _check_match(_head(pri), pri, tb, tc)


def test_match_ssd512_100gt_exact():
    pri = H.priors("ssd512")
    _, _, tb, tc = H.train_inputs(5, 3, pri.shape[0], min_gt=100, max_gt=100)
    _check_match(_head(pri), pri, tb, tc)


def test_match_ties_and_degenerate():
    pri = H.priors()
    box = torch.tensor([[0.2, 0.2, 0.6, 0.7]])
    tb = [
        torch.cat([box, box, box]),                                   # T3: identical gts claim one prior, last wins
        torch.tensor([[0.5, 0.5, 0.5, 0.5], [0.1, 0.1, 0.3, 0.3]]),   # zero-area gt: all IoU 0 -> prior 0 forced (T2)
        torch.tensor([[0.0, 0.0, 1.0, 1.0]]),
        torch.tensor([[0.3, 0.3, 0.31, 0.31], [0.3, 0.3, 0.31, 0.31], [0.9, 0.9, 1.0, 1.0]]),
    ]
    tc = [torch.tensor([3., 7., 5.]), torch.tensor([1., 2.]), torch.tensor([19.]), torch.tensor([0., 4., 8.])]
    _check_match(_head(pri), pri, tb, tc)


def _check_loss(pri, loc, conf, tb, tc, neg_ratio=3):
    head = _head(pri)
    gt = _packed(tb, tc, head)
    out = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True, taps=True, neg_ratio=neg_ratio)
    torch.cuda.synchronize()
    ref = O.multibox_loss(loc, conf, tb, tc, pri, ratio=neg_ratio)
    P = pri.shape[0]
    # positive count, CE values
    assert int(out["npos"][-1]) == ref["npos_total"]
    ce = out["ce"].cpu()
    assert torch.allclose(ce, ref["cce"], rtol=REL, atol=1e-6), f"CE max err {(ce - ref['cce']).abs().max()}"
    # mined set: exact under T4; where GPU and CPU CE differ in the last bits at the rank boundary the
    # disagreeing priors must sit on that boundary (|CE - threshold| within a few ulp)
    mined = H.unpack_mask(out["mined_mask"], P)
    diff = mined ^ ref["mined"]
    if diff.any():
        for b in diff.any(dim=1).nonzero().flatten().tolist():
            v = ref["cce"][b].clone()
            v[ref["pos"][b]] = 0
            k = int(3 * ref["npos"][b])
            thr = torch.sort(v, descending=True).values[min(k, P) - 1]
            bad = v[diff[b]]
            assert ((bad - thr).abs() <= 4e-6 * thr.abs().clamp(min=1)).all(), \
                f"image {b}: mined set differs away from the rank boundary"
        assert int(diff.sum()) <= 2 * loc.shape[0], "too many boundary disagreements"
    # the mined set has exactly k entries per image (when enough negatives exist)
    k = torch.minimum(neg_ratio * ref["npos"], (~ref["pos"]).sum(1))
    assert torch.equal(mined.sum(1), k)
    # losses
    losses = out["losses"].cpu()
    assert abs(losses[0].item() - ref["loc_loss"].item()) <= REL * abs(ref["loc_loss"].item()), (losses, ref["loc_loss"])
    assert abs(losses[1].item() - ref["conf_loss"].item()) <= REL * abs(ref["conf_loss"].item()), (losses, ref["conf_loss"])
    # gradients (closed form == autograd, pinned in test_oracle_vs_reference)
    ref_sel = dict(ref)
    ref_sel["mined"] = mined          # compare gradients on the GPU's own mined set
    gl, gc = O.multibox_grads(loc, conf, ref_sel)
    assert torch.allclose(out["grad_loc"].cpu(), gl, rtol=REL, atol=1e-9)
    gcd = out["grad_conf"].cpu()
    assert torch.allclose(gcd, gc, rtol=1e-4, atol=1e-8), f"grad_conf max err {(gcd - gc).abs().max()}"
    assert torch.equal(gcd.abs().sum(-1) != 0, (gc.abs().sum(-1) != 0)), "non-zero gradient rows"
    return out, ref


@pytest.mark.parametrize("seed,B", [(1, 8), (2, 32)])
def test_loss_ssd300(seed, B):
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(seed, B, pri.shape[0])
    _check_loss(pri, loc, conf, tb, tc)


def test_loss_ssd512_stress():
    pri = H.priors("ssd512")
    loc, conf, tb, tc = H.train_inputs(5, 3, pri.shape[0], min_gt=100, max_gt=100)
    _check_loss(pri, loc, conf, tb, tc)


def test_loss_unaligned_prior_count_uses_plain_copies():
    pri = H.priors()[:8731].contiguous()          # P % 4 != 0: no 16-byte aligned slices -> non-TMA path
    loc, conf, tb, tc = H.train_inputs(7, 4, pri.shape[0])
    _check_loss(pri, loc, conf, tb, tc)


def test_loss_small_prior_sets():
    for P in (4, 64, 1000, 2048):
        pri = H.priors()[5000:5000 + P].contiguous()
        loc, conf, tb, tc = H.train_inputs(11 + P, 3, P, max_gt=4)
        _check_loss(pri, loc, conf, tb, tc)


def test_mining_all_equal_ce_takes_lowest_indices():
    """T4: conf == 0 everywhere -> every CE equals log(21); the mined set is the 3*npos lowest-index negatives."""
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(3, 4, pri.shape[0])
    conf = torch.zeros_like(conf)
    out, ref = _check_loss(pri, loc, conf, tb, tc)
    mined = H.unpack_mask(out["mined_mask"], pri.shape[0])
    assert torch.equal(mined, ref["mined"])
    for b in range(4):
        neg = (~ref["pos"][b]).nonzero().flatten()
        k = int(3 * ref["npos"][b])
        assert torch.equal(mined[b].nonzero().flatten(), neg[:k])


def test_mining_partial_ties():
    """Ties that straddle the rank boundary inside otherwise distinct values."""
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(4, 4, pri.shape[0])
    conf[:, 1000:3000:7] = 0.0          # blocks of identical rows -> identical CE among negatives
    conf[:, 1000:3000:7, 20] = -3.0     # make them large-loss negatives so they sit near the top
    out, ref = _check_loss(pri, loc, conf, tb, tc)
    assert torch.equal(H.unpack_mask(out["mined_mask"], pri.shape[0]), ref["mined"])


def test_neg_ratio_saturates():
    """k >= P: every negative is mined (Losses.py:194 takes the whole row)."""
    pri = H.priors()[:2048].contiguous()
    loc, conf, tb, tc = H.train_inputs(9, 2, 2048)
    _check_loss(pri, loc, conf, tb, tc, neg_ratio=5000)


def test_forward_only_matches_fused():
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(6, 8, pri.shape[0])
    head = _head(pri)
    gt = _packed(tb, tc, head)
    a = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=False)
    b = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True)
    torch.cuda.synchronize()
    assert torch.equal(a["losses"], b["losses"]) and torch.equal(a["sums"], b["sums"])
    assert a["grad_loc"] is None


def test_loss_is_run_to_run_deterministic():
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(8, 16, pri.shape[0])
    head = _head(pri)
    gt = _packed(tb, tc, head)
    l, c = loc.cuda(), conf.cuda()
    outs = [head.loss(l, c, gt, with_grads=True) for _ in range(3)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o["sums"], outs[0]["sums"])
        assert torch.equal(o["grad_conf"], outs[0]["grad_conf"])
        assert torch.equal(o["grad_loc"], outs[0]["grad_loc"])


def test_autograd_surface():
    """ssd()-style use: two scalars with grad_fn; backward with non-unit upstream gradients."""
    from objectdetection_ssd_b200.head import multibox_loss
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(10, 4, pri.shape[0])
    head = _head(pri)
    l = loc.cuda().requires_grad_(True)
    c = conf.cuda().requires_grad_(True)
    l1, l2 = multibox_loss(head, l, c, [b.cuda() for b in tb], [x.cuda() for x in tc])
    (2.0 * l1 + 0.5 * l2).backward()
    ref = O.multibox_loss(loc, conf, tb, tc, pri)
    gl, gc = O.multibox_grads(loc, conf, ref, 2.0, 0.5)
    assert torch.allclose(l.grad.cpu(), gl, rtol=REL, atol=1e-9)
    assert torch.allclose(c.grad.cpu(), gc, rtol=1e-4, atol=1e-8)
    assert abs(l1.item() - ref["loc_loss"].item()) <= REL * abs(ref["loc_loss"].item())


def test_image_without_gt_raises_like_reference():
    from objectdetection_ssd_b200.head import multibox_loss
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(10, 2, pri.shape[0])
    tb[1] = torch.zeros(0, 4)
    tc[1] = torch.zeros(0)
    with pytest.raises(IndexError):
        multibox_loss(_head(pri), loc.cuda(), conf.cuda(), tb, tc)


def test_loss_misaligned_conf_pointer_uses_plain_loads():
    """A conf view that is not 16-byte aligned cannot use bulk copies: the kernel falls back to plain loads."""
    pri = H.priors()
    loc, conf, tb, tc = H.train_inputs(12, 3, pri.shape[0])
    head = _head(pri)
    gt = _packed(tb, tc, head)
    big = torch.empty(conf.numel() + 1, device="cuda")
    view = big[1:].view_as(conf)
    view.copy_(conf)
    assert view.data_ptr() % 16 != 0
    a = head.loss(loc.cuda(), view, gt, with_grads=True)
    b = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True)
    torch.cuda.synchronize()
    assert torch.equal(a["losses"], b["losses"])
    assert torch.equal(a["grad_conf"], b["grad_conf"]) and torch.equal(a["grad_loc"], b["grad_loc"])


def test_fused_match_equals_standalone_match():
    """The match fused into the CE streaming kernel must reproduce ssdhead_match bit for bit (class bytes,
    best prior per gt, positive counts), including the tie cases and a warp that straddles two images."""
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    pri = H.priors()
    for seed, B, gts in ((51, 5, None), (52, 3, "ties")):
        loc, conf, tb, tc = H.train_inputs(seed, B, pri.shape[0])
        if gts == "ties":
            box = torch.tensor([[0.2, 0.2, 0.6, 0.7]])
            tb = [torch.cat([box, box, box]), torch.tensor([[0.5, 0.5, 0.5, 0.5], [0.1, 0.1, 0.3, 0.3]]),
                  torch.tensor([[0.3, 0.3, 0.31, 0.31], [0.3, 0.3, 0.31, 0.31], [0.9, 0.9, 1.0, 1.0]])]
            tc = [torch.tensor([3., 7., 5.]), torch.tensor([1., 2.]), torch.tensor([0., 4., 8.])]
        head = MultiboxHead(pri, "cuda")
        gt = PackedGT(tb, tc, head.dev)
        ref = head.match(gt)
        out = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True)
        torch.cuda.synchronize()
        assert torch.equal(out["cls_u8"], ref["cls_u8"])
        assert torch.equal(out["best_prior"][:gt.sumG], ref["best_prior"][:gt.sumG])
        assert torch.equal(out["npos"], ref["npos"])
        again = head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True, match=ref)      # separate-match route
        torch.cuda.synchronize()
        assert torch.equal(again["losses"], out["losses"]) and torch.equal(again["grad_conf"], out["grad_conf"])


def test_mining_stage_isolated_is_bit_exact():
    """Stage-isolated T4 check: feed the ORACLE's cross entropies to ssdhead_mine (skipping the kernels' own exp/log)
    and require the mined-negative set to equal the oracle's bit for bit, for random and tie-heavy inputs."""
    from objectdetection_ssd_b200 import _lib
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    pri = H.priors()
    P = pri.shape[0]
    lib = _lib.load()
    for seed, B, quant in ((61, 8, None), (62, 4, 0.25)):
        loc, conf, tb, tc = H.train_inputs(seed, B, P)
        if quant:
            conf = torch.round(conf / quant) * quant          # many exactly equal CE values -> ties across the boundary
        ref = O.multibox_loss(loc, conf, tb, tc, pri)
        head = MultiboxHead(pri, "cuda")
        gt = PackedGT(tb, tc, head.dev)
        m = head.match(gt)
        bg = torch.full((B, P), 20, dtype=torch.int64)
        ce_bg = O.cross_entropy_rows(conf, bg).cuda().contiguous()        # what the streaming kernel would hand over
        l, c = loc.cuda(), conf.cuda()
        sums = torch.empty(2, dtype=torch.float64, device="cuda")
        losses = torch.empty(2, device="cuda")
        mined = torch.empty(B, (P + 31) // 32, dtype=torch.int32, device="cuda")
        ws = head._workspace(_lib.WS_LOSS, B, 0)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.ssdhead_mine(l.data_ptr(), c.data_ptr(), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
                                    head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), m["best_prior"].data_ptr(),
                                    m["npos"].data_ptr(), m["npos"][B:].data_ptr(), m["cls_u8"].data_ptr(),
                                    B, P, 21, 3, 0.5, sums.data_ptr(), losses.data_ptr(), None, None,
                                    mined.data_ptr(), ce_bg.data_ptr(), ws.data_ptr(), ws.numel(), st), "ssdhead_mine")
        torch.cuda.synchronize()
        assert torch.equal(H.unpack_mask(mined, P), ref["mined"]), f"seed {seed}: mined set differs from the oracle"
        assert abs(losses[1].item() - ref["conf_loss"].item()) <= 1e-6 * ref["conf_loss"].item()


@pytest.mark.parametrize("counts", [(150, 3, 70), (129, 1), (65, 64, 33)])
def test_many_gt_per_image_slow_paths(counts):
    """Images with more gts than the shared-memory staging capacities (64 / 128) and the 32-gt shuffle chunks:
    exercises the multi-chunk paths of match_kernel, the fused match, the finaliser(s) and mine_kernel."""
    from objectdetection_ssd_b200 import synth
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    pri = H.priors()
    P = pri.shape[0]
    tb, tc = [], []
    for i, n in enumerate(counts):
        gb, gc = synth.make_gt(90 + i, 1, n, n)
        tb.append(torch.from_numpy(gb[0]))
        tc.append(torch.from_numpy(gc[0]))
    B = len(counts)
    loc, conf = synth.make_head(91, B, P)
    loc, conf = torch.from_numpy(loc), torch.from_numpy(conf)
    _check_match(_head(pri), pri, tb, tc)                              # standalone match kernel
    out, ref = _check_loss(pri, loc, conf, tb, tc)                     # two-kernel step (fused match + fused finaliser)
    assert torch.equal(out["cls_u8"].cpu().long(), ref["cls"])
    assert torch.equal(out["best_prior"][:sum(counts)].cpu().long(), ref["best_prior"])
    # three-kernel route (what a sharded batch uses): ce_match_stream + finaliser kernel + plain mining kernel
    from objectdetection_ssd_b200 import _lib
    head = MultiboxHead(pri, "cuda")
    gt = PackedGT(tb, tc, head.dev)
    lib = _lib.load()
    m = head._match_outputs(gt, False)
    ws = head._workspace(_lib.WS_LOSS, B, 0)
    wm = head._workspace(_lib.WS_MATCH, B, gt.sumG)
    st = torch.cuda.current_stream().cuda_stream
    l, c = loc.cuda(), conf.cuda()
    gl, gcf = torch.empty_like(l), torch.empty_like(c)
    sums = torch.empty(2, dtype=torch.float64, device="cuda")
    losses = torch.empty(2, device="cuda")
    _lib.check(lib.ssdhead_ce_match_stream(c.data_ptr(), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
                                           head.pri_xyxy.data_ptr(), B, P, 21, gt.sumG, 0.5, None, gl.data_ptr(), gcf.data_ptr(),
                                           m["cls_u8"].data_ptr(), m["best_prior"].data_ptr(), m["npos"].data_ptr(),
                                           ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), 1, st), "ce_match_stream")
    _lib.check(lib.ssdhead_mine(l.data_ptr(), c.data_ptr(), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
                                head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), m["best_prior"].data_ptr(),
                                m["npos"].data_ptr(), m["npos"][B:].data_ptr(), m["cls_u8"].data_ptr(), B, P, 21, 3, 0.5,
                                sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), None, None,
                                ws.data_ptr(), ws.numel(), st), "mine")
    torch.cuda.synchronize()
    assert torch.equal(m["cls_u8"], out["cls_u8"]) and torch.equal(m["npos"], out["npos"])
    assert torch.equal(gcf, out["grad_conf"]) and torch.equal(gl, out["grad_loc"])
    assert torch.equal(losses, out["losses"])


def test_seeded_cull_of_the_fused_match_changes_nothing():
    """Batches with many gts per image skip, per warp, the gts whose area ratio with the warp's priors can reach neither
    pos_iou nor a sampled lower bound of the gt's best IoU (csrc/loss.cu, match_seed_kernel).  With the cull forced on
    and forced off the step must produce the same bits: class map, best priors, positive counts, losses, gradients -
    for an ordinary batch, for many gts per image (tiny, huge and degenerate boxes included), and for ties."""
    import os
    from objectdetection_ssd_b200 import synth
    from objectdetection_ssd_b200.head import PackedGT
    pri = H.priors()
    P = pri.shape[0]
    head = _head(pri)
    cases = []
    loc, conf, tb, tc = H.train_inputs(95, 6, P)                                   # ordinary: 1..10 gts per image
    cases.append((loc, conf, tb, tc))
    gb, gc = synth.make_gt(96, 3, 60, 100)                                         # many gts per image
    tb = [torch.from_numpy(b).clone() for b in gb]
    tc = [torch.from_numpy(c) for c in gc]
    tb[0][0] = torch.tensor([0.50, 0.50, 0.505, 0.505])                            # tiny
    tb[0][1] = torch.tensor([0.0, 0.0, 1.0, 1.0])                                  # the whole image
    tb[0][2] = torch.tensor([0.3, 0.3, 0.3001, 0.6])                               # a sliver
    tb[1][0] = tb[1][1].clone()                                                    # two identical gts (T2 / T3 ties)
    tb[2][5] = torch.cat([pri[4000, :2] - pri[4000, 2:] / 2, pri[4000, :2] + pri[4000, 2:] / 2])   # exactly a prior
    l2, c2 = synth.make_head(97, 3, P)
    cases.append((torch.from_numpy(l2), torch.from_numpy(c2), tb, tc))
    old = os.environ.get("SSDHEAD_MATCH_CULL_MIN")
    try:
        for loc, conf, tb, tc in cases:
            outs = []
            for setting in ("1", "1000000"):
                os.environ["SSDHEAD_MATCH_CULL_MIN"] = setting
                gt = PackedGT(tb, tc, head.dev)
                outs.append(head.loss(loc.cuda(), conf.cuda(), gt, with_grads=True))
                torch.cuda.synchronize()
            a, b = outs
            n = sum(int(x.shape[0]) for x in tb)
            assert torch.equal(a["cls_u8"], b["cls_u8"]) and torch.equal(a["npos"], b["npos"])
            assert torch.equal(a["best_prior"][:n], b["best_prior"][:n])
            assert torch.equal(a["losses"], b["losses"]) and torch.equal(a["sums"], b["sums"])
            assert torch.equal(a["grad_loc"], b["grad_loc"]) and torch.equal(a["grad_conf"], b["grad_conf"])
        ref = O.multibox_loss(cases[1][0], cases[1][1], cases[1][2], cases[1][3], pri)      # and the oracle agrees
        assert torch.equal(a["cls_u8"].cpu().long(), ref["cls"])
        assert torch.equal(a["best_prior"][:n].cpu().long(), ref["best_prior"])
    finally:
        if old is None:
            os.environ.pop("SSDHEAD_MATCH_CULL_MIN", None)
        else:
            os.environ["SSDHEAD_MATCH_CULL_MIN"] = old
