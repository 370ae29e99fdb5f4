"""CPU: the native gt collate (ssdhead_pack_gt) against the oracle's restatement of
Dataset.py:28-36 + Losses.py:129-130 - bit-exact (plain fp32 copies and divisions)."""
import numpy as np
import pytest
import torch

from oracle import ssd_oracle as O


def _ragged(seed, B, lo=1, hi=12, pixels=False):
    g = np.random.default_rng(seed)
    boxes, classes, diff, wh = [], [], [], []
    for _ in range(B):
        n = int(g.integers(lo, hi + 1))
        w, h = float(g.integers(200, 640)), float(g.integers(200, 640))
        xy = g.uniform(0, 0.7, (n, 2)).astype(np.float32)
        sz = g.uniform(0.05, 0.3, (n, 2)).astype(np.float32)
        b = np.concatenate([xy, xy + sz], 1)
        if pixels:
            b = (b * np.array([w, h, w, h], np.float32)).round()
        boxes.append(torch.from_numpy(b.astype(np.float32)))
        classes.append(torch.from_numpy(g.integers(0, 20, n).astype(np.float32)))
        d = (g.uniform(size=n) < 0.3).astype(np.uint8)
        d[int(g.integers(0, n))] = 0                      # at least one easy box per image
        diff.append(torch.from_numpy(d))
        wh.append((w, h))
    return boxes, classes, diff, np.array(wh, np.float32)


@pytest.mark.parametrize("keep_difficult,pixels", [(True, False), (False, False), (False, True), (True, True)])
def test_collate_matches_the_reference_ops(keep_difficult, pixels):
    from objectdetection_ssd_b200.collate import collate_gt
    boxes, classes, diff, wh = _ragged(5, 33, pixels=pixels)
    rb, rc, ro = O.collate_gt(boxes, classes, diff, keep_difficult, wh if pixels else None)
    gb, gc, go = collate_gt(boxes, classes, diff, keep_difficult, wh if pixels else None)
    assert np.array_equal(go, ro.numpy())
    assert np.array_equal(gb.view(np.uint32), rb.numpy().view(np.uint32))           # same bits, incl. the fp32 division
    assert np.array_equal(gc, rc.numpy())


def test_collate_accepts_numpy_lists_and_int_classes():
    from objectdetection_ssd_b200.collate import collate_gt
    boxes, classes, _, _ = _ragged(6, 4)
    gb, gc, go = collate_gt([b.numpy() for b in boxes], [c.numpy().astype(np.int64) for c in classes])
    rb, rc, ro = O.collate_gt(boxes, classes)
    assert np.array_equal(gb, rb.numpy()) and np.array_equal(gc, rc.numpy()) and np.array_equal(go, ro.numpy())


def test_collate_image_without_boxes_raises_like_the_reference():
    from objectdetection_ssd_b200.collate import collate_gt
    boxes, classes, diff, _ = _ragged(7, 3)
    boxes[1] = torch.zeros(0, 4)
    classes[1] = torch.zeros(0)
    with pytest.raises(IndexError):
        collate_gt(boxes, classes)
    boxes, classes, diff, _ = _ragged(8, 3)
    diff[2][:] = 1                                        # every box of image 2 is difficult
    with pytest.raises(IndexError):
        collate_gt(boxes, classes, diff, keep_difficult=False)
    assert collate_gt(boxes, classes, diff, keep_difficult=True)[2][-1] == sum(b.shape[0] for b in boxes)


def test_collate_feeds_the_packed_layout_of_synth():
    from objectdetection_ssd_b200 import synth
    from objectdetection_ssd_b200.collate import collate_gt
    gb, gc = synth.make_gt(3, 16)
    a = synth.pack_gt(gb, gc)
    b = collate_gt(gb, gc, pinned=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_pack_gt_c_entry_error_codes():
    """The C entry point directly: capacity check, bad arguments, empty batch."""
    import ctypes as C
    from objectdetection_ssd_b200 import _lib
    lib = _lib.load()
    b0 = np.array([[0.1, 0.1, 0.5, 0.5], [0.2, 0.2, 0.9, 0.9]], np.float32)
    c0 = np.array([3, 7], np.float32)
    boxes = (C.c_void_p * 1)(b0.ctypes.data)
    classes = (C.c_void_p * 1)(c0.ctypes.data)
    counts = np.array([2], np.int32)
    ob, oc, oo = np.zeros((2, 4), np.float32), np.zeros(2, np.float32), np.zeros(2, np.int32)
    call = lambda cap, bx=boxes, out_off=oo.ctypes.data: lib.ssdhead_pack_gt(
        bx, classes, None, counts.ctypes.data, 1, 1, None, ob.ctypes.data, oc.ctypes.data, out_off, cap)
    assert call(2) == 2 and np.array_equal(ob, b0) and np.array_equal(oo, [0, 2])
    assert call(1) == _lib.E_WORKSPACE                      # room for one box only
    assert call(2, out_off=None) == _lib.E_BADARG
    assert call(2, bx=None) == _lib.E_BADARG
    assert lib.ssdhead_pack_gt(None, None, None, None, 0, 1, None, None, None, oo.ctypes.data, 0) == 0   # empty batch


def test_collate_property_random_ragged_batches():
    """Hypothesis: any ragged batch (1..20 images, 1..30 boxes, random difficult flags with at least one easy box) packs
    exactly as the reference's ops do."""
    from hypothesis import given, settings, strategies as st
    from objectdetection_ssd_b200.collate import collate_gt

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.integers(1, 20), st.booleans(), st.booleans())
    def check(seed, B, keep_difficult, pixels):
        boxes, classes, diff, wh = _ragged(seed, B, lo=1, hi=30, pixels=pixels)
        rb, rc, ro = O.collate_gt(boxes, classes, diff, keep_difficult, wh if pixels else None)
        gb, gc, go = collate_gt(boxes, classes, diff, keep_difficult, wh if pixels else None, pinned=False)
        assert np.array_equal(go, ro.numpy()) and np.array_equal(gc, rc.numpy())
        assert np.array_equal(gb.view(np.uint32), rb.numpy().view(np.uint32))

    check()


def test_collate_result_survives_unpacking_and_gc():
    """the arrays must keep their (page-locked) allocation alive on their own - unpack the result,
    drop everything else, collect, allocate more of the same kind, and the contents must still be there."""
    import gc
    from objectdetection_ssd_b200.collate import collate_gt
    boxes, classes, diff, wh = _ragged(11, 17)
    rb, rc, ro = O.collate_gt(boxes, classes, diff, True, None)
    gb, gc_, go = collate_gt(boxes, classes, diff, True, None)          # tuple dropped at once
    gc.collect()
    junk = [collate_gt(boxes, classes, diff, True, None) for _ in range(8)]   # would reuse a freed block
    for j in junk:
        j[0][...] = -1.0
    del junk
    gc.collect()
    assert np.array_equal(gb, rb.numpy()) and np.array_equal(gc_, rc.numpy()) and np.array_equal(go, ro.numpy())
    tail = gb[3:]                                                        # a slice alone keeps the block alive too
    del gb, gc_, go
    gc.collect()
    assert np.array_equal(tail, rb.numpy()[3:])


def test_gt_staging_layout_is_16_byte_aligned():
    from objectdetection_ssd_b200.pinned import gt_layout
    for cap, B in ((1, 1), (63, 7), (2048, 256), (12800, 128)):
        o0, o1, o2, total = gt_layout(cap, B)
        assert o0 == 0 and o1 % 16 == 0 and o2 % 16 == 0 and total % 16 == 0
        assert o1 >= cap * 16 and o2 - o1 >= cap * 4 and total - o2 >= (B + 1) * 4
