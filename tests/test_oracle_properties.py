"""CPU property tests (hypothesis) of the oracle: the invariants the GPU parity tests lean on."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ssd_oracle as O

PRI = O.make_priors()
PXY = O.cxcywh_to_xyxy(PRI)


def boxes_strategy(max_n=8):
    coord = st.floats(0.0, 1.0, allow_nan=False, width=32)
    box = st.tuples(coord, coord, coord, coord).map(
        lambda t: (min(t[0], t[2]), min(t[1], t[3]), max(t[0], t[2]), max(t[1], t[3])))
    return st.lists(box, min_size=1, max_size=max_n).map(lambda l: torch.tensor(l, dtype=torch.float32))


@settings(max_examples=25, deadline=None)
@given(boxes_strategy(), boxes_strategy())
def test_iou_is_symmetric_bounded_and_exact_on_itself(a, b):
    iou = O.iou_matrix(a, b)
    # a+b == b+a, min/max commute: bit-symmetric (two zero-area boxes give 0/0 = nan on both sides)
    assert torch.equal(torch.nan_to_num(iou, nan=-1.0), torch.nan_to_num(O.iou_matrix(b, a).t(), nan=-1.0))
    ok = ~torch.isnan(iou)                                               # nan only for two zero-area boxes (0/0)
    assert ((iou[ok] >= 0) & (iou[ok] <= 1.0000001)).all()
    area = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    self_iou = O.iou_matrix(a, a).diagonal()
    assert torch.equal(self_iou[area > 0], torch.ones(int((area > 0).sum())))


@settings(max_examples=15, deadline=None)
@given(boxes_strategy(6), st.integers(0, 2 ** 31 - 1))
def test_match_invariants(gt, seed):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randint(0, 20, (gt.shape[0],), generator=g).float()
    iou = O.iou_matrix(gt, PXY)
    c, obj, overlap, best_prior = O.match_image(iou, cls)
    # every gt's best prior ends up positive and owned by the LAST gt that claimed it (T3)
    for gi in range(gt.shape[0]):
        p = int(best_prior[gi])
        owner = max(j for j in range(gt.shape[0]) if int(best_prior[j]) == p)
        assert int(obj[p]) == owner and float(overlap[p]) == 1.0 and float(c[p]) == float(cls[owner])
    # non-forced priors: positive <=> not (max IoU < 0.5), class of the first arg-max gt (T1)
    forced = torch.zeros(PXY.shape[0], dtype=torch.bool)
    forced[best_prior] = True
    mx, am = iou.max(dim=0)
    nf = ~forced
    assert torch.equal(obj[nf], am[nf])
    assert torch.equal(c[nf] != 20, ~(mx[nf] < 0.5))
    assert 1 <= int((c != 20).sum())


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.floats(0.2, 0.7))
def test_nms_keep_set_properties(seed, n, thr):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g)
    s = torch.rand(n, 2, generator=g) * 0.4 + 0.05
    boxes = torch.cat([c - s / 2, c + s / 2], 1)
    keep = O.nms_sorted(boxes, thr)
    assert bool(keep[0])                                                 # the best-scored box always survives
    kept = boxes[keep]
    iou = O.iou_matrix(kept, kept)
    iou.fill_diagonal_(0)
    assert (iou < thr).all()                                             # survivors do not suppress each other
    sup = boxes[~keep]
    if sup.shape[0]:
        assert (O.iou_matrix(sup, kept) >= thr).any(dim=1).all()         # every dropped box is covered by a survivor
    assert torch.equal(O.nms_sorted(kept, thr), torch.ones(kept.shape[0], dtype=torch.bool))    # idempotent


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 5))
def test_mining_is_the_top_k_under_rule_t4(seed, npos_per_row):
    g = torch.Generator().manual_seed(seed)
    P = 300
    cce = torch.rand(2, P, generator=g)
    cce[:, ::7] = 0.5                                                    # plenty of exact ties
    pos = torch.zeros(2, P, dtype=torch.bool)
    for r in range(2):
        pos[r, torch.randperm(P, generator=g)[:npos_per_row]] = True
    mined = O.mine_hard_negatives(cce, pos, 3)
    for r in range(2):
        k = 3 * npos_per_row
        v = cce[r].clone()
        v[pos[r]] = 0
        order = sorted(range(P), key=lambda j: (-float(v[j]), j))[:k]    # descending value, ties -> lower index
        want = torch.zeros(P, dtype=torch.bool)
        want[order] = True
        want &= ~pos[r]
        assert torch.equal(mined[r], want)


def test_encode_decode_round_trip():
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, PRI.shape[0], (500,), generator=g)
    c = torch.rand(500, 2, generator=g) * 0.8 + 0.1
    s = torch.rand(500, 2, generator=g) * 0.5 + 0.02
    box = torch.cat([c, s], 1)
    back = O.decode(O.encode(box, PRI[idx]), PRI[idx])
    assert torch.allclose(back, box, rtol=1e-5, atol=1e-6)
    assert torch.allclose(O.xyxy_to_cxcywh(O.cxcywh_to_xyxy(box)), box, rtol=1e-6, atol=1e-7)
