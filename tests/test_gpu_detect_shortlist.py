"""GPU: the two routes of the detect kernels give the same output, bit for bit.

Short-list route (default): a sampled score floor, only candidates above it are listed, the sweep runs on that short
list; an image it cannot decide is listed again in full by its sweep CTA.  ``SSDHEAD_DETECT_SHORTLIST=0`` runs the
exhaustive kernels (every candidate >= min_score listed up front) for every image.  The tests compare the two routes
with each other and with the oracle, and check with ``ssdhead_detect_fallbacks`` that each case really took the route
it is meant to cover.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import ssd_oracle as O
from tests import helpers as H
from tests.test_gpu_detect import _check_exact, _oracle_stage

pytestmark = pytest.mark.gpu

KEYS = ("boxes", "prob", "cls", "prior")


class _route:
    def __init__(self, shortlist):
        self.v = "1" if shortlist else "0"

    def __enter__(self):
        self.old = os.environ.get("SSDHEAD_DETECT_SHORTLIST")
        os.environ["SSDHEAD_DETECT_SHORTLIST"] = self.v

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("SSDHEAD_DETECT_SHORTLIST", None)
        else:
            os.environ["SSDHEAD_DETECT_SHORTLIST"] = self.old


def _same(a, b):
    ca, cb = a["cnt"].cpu(), b["cnt"].cpu()
    assert torch.equal(ca, cb), (ca.tolist(), cb.tolist())
    for i, k in enumerate(ca.tolist()):
        for key in KEYS:
            assert torch.equal(a[key][i, :k].cpu(), b[key][i, :k].cpu()), f"image {i}: {key} differs between the routes"


def _few_sites_inputs(seed, B, P, classes, sites=40):
    """Every prior's box sits on one of ``sites`` places (tiny jitter: IoU ~ 1 inside a site), scores uniform in
    ``classes``: thousands of candidates, at most ``sites`` survivors per class -> far fewer than top_k + 1 among the best
    ~1600, so the short list cannot decide the image."""
    g = torch.Generator().manual_seed(seed)
    centres = 0.15 + 0.7 * torch.rand(B, sites, 2, generator=g)
    which = torch.randint(0, sites, (B, P), generator=g)
    boxes = torch.zeros(B, P, 4)
    boxes[..., :2] = torch.gather(centres, 1, which.unsqueeze(-1).expand(B, P, 2)) + 1e-4 * torch.randn(B, P, 2, generator=g)
    boxes[..., 2:] = 0.05
    probs = torch.zeros(B, P, 21)
    for c in classes:
        probs[..., c] = 0.05 + 0.4 * torch.rand(B, P, generator=g)
    probs[..., 20] = 1.0 - probs[..., :20].sum(-1)
    return boxes.contiguous(), probs.contiguous()


def _head(pri):
    from objectdetection_ssd_b200.head import MultiboxHead
    return MultiboxHead(pri, "cuda")


def test_shortlist_decides_ordinary_images_and_equals_the_exhaustive_route():
    from objectdetection_ssd_b200.head import detect, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    for seed, B, bias, min_score, top_k in ((51, 6, 6.0, 0.01, 200), (52, 3, 8.0, 0.01, 200), (53, 2, 4.0, 0.01, 50)):
        loc, conf = H.detect_inputs(seed, B, pri.shape[0], bg_bias=bias)
        with _route(True):
            fast = detect(head, loc, conf, min_score, 0.45, top_k)
            nfb = detect_fallbacks(head, B)
        with _route(False):
            full = detect(head, loc, conf, min_score, 0.45, top_k)
            nfb0 = detect_fallbacks(head, B)
        torch.cuda.synchronize()
        _same(fast, full)
        assert nfb == 0, f"seed {seed}: {nfb} of {B} ordinary images were not decided by their short list"
        assert nfb0 == 0                                         # the exhaustive route flags nothing
        assert (fast["cnt"].cpu() == top_k).all()


def test_heavy_suppression_is_flagged_and_served_by_the_exhaustive_kernels():
    """8732 candidates on 40 sites in each of two classes: the ~1600 best keep at most 80 boxes, so the short list cannot
    decide; the image is flagged and the result must still be the oracle's."""
    from objectdetection_ssd_b200.head import detect_from_scores, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    boxes, probs = _few_sites_inputs(54, 3, pri.shape[0], classes=(2, 11))
    with _route(True):
        out = detect_from_scores(head, boxes, probs, 0.05, 0.45, 200)
        nfb = detect_fallbacks(head, 3)
    torch.cuda.synchronize()
    assert nfb == 3, nfb
    _check_exact(out, _oracle_stage(boxes, probs, 0.05, 0.45, 200), 200)
    # top_k small enough for the short list to decide: nothing is flagged, same answer as the oracle
    with _route(True):
        out = detect_from_scores(head, boxes, probs, 0.05, 0.45, 7)
        nfb = detect_fallbacks(head, 3)
    assert nfb == 0, nfb
    _check_exact(out, _oracle_stage(boxes, probs, 0.05, 0.45, 7), 7)
    # the count describes the LAST call: a call on the exhaustive route after a call that needed the fallback reports 0
    with _route(True):
        detect_from_scores(head, boxes, probs, 0.05, 0.45, 200)
        assert detect_fallbacks(head, 3) == 3
    with _route(False):
        out = detect_from_scores(head, boxes, probs, 0.05, 0.45, 200)
        assert detect_fallbacks(head, 3) == 0
    _check_exact(out, _oracle_stage(boxes, probs, 0.05, 0.45, 200), 200)


def test_mixed_batch_only_the_undecided_images_take_the_fallback():
    """Undecided images interleaved with ordinary ones: only they are listed twice."""
    from objectdetection_ssd_b200.head import detect_from_scores, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    B = 80
    loc, conf = H.detect_inputs(55, B, pri.shape[0], bg_bias=6.0)
    boxes = torch.stack([O.decode(loc[i], pri) for i in range(B)])
    probs = F.softmax(conf, dim=2)
    cb, cp = _few_sites_inputs(56, 4, pri.shape[0], classes=(5,))
    hard = [i for i in range(B) if i % 2 == 1]                   # 40 images
    for j, i in enumerate(hard):
        boxes[i], probs[i] = cb[j % 4], cp[j % 4]
    with _route(True):
        fast = detect_from_scores(head, boxes, probs, 0.02, 0.45, 200)
        nfb = detect_fallbacks(head, B)
    with _route(False):
        full = detect_from_scores(head, boxes, probs, 0.02, 0.45, 200)
    torch.cuda.synchronize()
    assert nfb == len(hard), (nfb, len(hard))
    _same(fast, full)
    for i in (0, 1, 2, 3, 78, 79):
        _check_exact({k: v[i:i + 1] for k, v in fast.items()}, _oracle_stage(boxes[i:i + 1], probs[i:i + 1], 0.02, 0.45, 200), 200)


def test_repeated_calls_leave_the_workspace_clean():
    """Flagged and ordinary batches alternate on one workspace: counters, histogram and flag list must reset themselves."""
    from objectdetection_ssd_b200.head import detect_from_scores, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    loc, conf = H.detect_inputs(57, 4, pri.shape[0], bg_bias=6.0)
    boxes = torch.stack([O.decode(loc[i], pri) for i in range(4)])
    probs = F.softmax(conf, dim=2)
    cb, cp = _few_sites_inputs(58, 4, pri.shape[0], classes=(2, 11))
    ref_a = _oracle_stage(boxes, probs, 0.01, 0.45, 200)
    ref_b = _oracle_stage(cb, cp, 0.05, 0.45, 200)
    with _route(True):
        for it in range(3):
            out = detect_from_scores(head, boxes, probs, 0.01, 0.45, 200)
            assert detect_fallbacks(head, 4) == 0
            _check_exact(out, ref_a, 200)
            out = detect_from_scores(head, cb, cp, 0.05, 0.45, 200)
            assert detect_fallbacks(head, 4) == 4
            _check_exact(out, ref_b, 200)


def test_few_candidates_keep_the_floor_at_min_score():
    """Fewer candidates than the sampling target: the floor stays at min_score, the short list IS the full list and an
    image with fewer than top_k survivors is decided (class-major output) without the fallback."""
    from objectdetection_ssd_b200.head import detect_from_scores, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    loc, conf = H.detect_inputs(32, 3, pri.shape[0], bg_bias=9.0)
    boxes = torch.stack([O.decode(loc[i], pri) for i in range(3)])
    probs = F.softmax(conf, dim=2)
    ref = _oracle_stage(boxes, probs, 0.05, 0.45, 200)
    assert all(0 < r[0].shape[0] <= 200 for r in ref)
    with _route(True):
        out = detect_from_scores(head, boxes, probs, 0.05, 0.45, 200)
        nfb = detect_fallbacks(head, 3)
    assert nfb == 0, nfb
    _check_exact(out, ref, 200)


def test_levels_entry_takes_the_same_routes():
    from objectdetection_ssd_b200.head import detect, detect_levels, detect_fallbacks
    pri = H.priors()
    head = _head(pri)
    B = 3
    loc, conf = H.detect_inputs(60, B, pri.shape[0], bg_bias=6.0)
    counts = [5776, 2166, 600, 150, 36, 4]                       # SSD300: 38^2*4, 19^2*6, 10^2*6, 5^2*6, 3^2*4, 1*4 (Model.py:212-235)
    locs, confs, s = [], [], 0
    for n in counts:
        locs.append(loc[:, s:s + n].contiguous().cuda())
        confs.append(conf[:, s:s + n].contiguous().cuda())
        s += n
    with _route(True):
        a = detect_levels(head, locs, confs, 0.01, 0.45, 200)
        nfb = detect_fallbacks(head, B)
        b = detect(head, loc, conf, 0.01, 0.45, 200)
    with _route(False):
        c = detect_levels(head, locs, confs, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    assert nfb == 0
    _same(a, b)
    _same(a, c)


def test_the_stage_isolated_suite_under_the_forced_short_list_route():
    """tests/test_gpu_detect.py runs its small batches on the exhaustive route (the automatic choice below ~900 k rows);
    here the same exact-equality tests - ties, heavy suppression through every slice, random shapes, unaligned prior
    counts, suppression chains - run with the short-list route forced."""
    import tests.test_gpu_detect as D
    with _route(True):
        for args in ((8.0, 0.01, 4), (6.0, 0.01, 2), (6.0, 0.2, 3), (2.0, 0.05, 2)):
            D.test_detect_from_scores_exact(*args)
        D.test_detect_fewer_than_topk_is_class_major_unsorted()
        D.test_detect_no_candidates_and_ties()
        D.test_detect_small_topk_and_pixel_scale()
        D.test_detect_ssd512_priors()
        D.test_detect_end_to_end_close()
        D.test_detect_heavy_suppression_runs_through_every_slice()
        D.test_detect_quantised_scores_tie_across_classes_and_priors()
        D.test_detect_unaligned_prior_count_takes_the_plain_load_path()
        D.test_detect_randomised_differential()
        D.test_detect_suppression_chains_and_sparse_overlaps()


def test_automatic_route_choice_by_call_size():
    """Unset switch: small calls take the exhaustive route (which can report an overflowed candidate cap), large calls the
    short list; both equal the forced routes."""
    from objectdetection_ssd_b200.head import detect
    pri = H.priors()
    head = _head(pri)
    old = os.environ.pop("SSDHEAD_DETECT_SHORTLIST", None)
    try:
        for B in (3, 120):
            loc, conf = H.detect_inputs(61, B, pri.shape[0], bg_bias=6.0)
            auto = detect(head, loc, conf, 0.01, 0.45, 200)
            with _route(B < 100):
                other = detect(head, loc, conf, 0.01, 0.45, 200)
            torch.cuda.synchronize()
            _same(auto, other)
    finally:
        if old is not None:
            os.environ["SSDHEAD_DETECT_SHORTLIST"] = old
