"""GPU: the one-call context front end (host buffers / device tensors) gives the same results as the
step-by-step C ABI and matches the oracle."""
import numpy as np
import pytest
import torch

from objectdetection_ssd_b200 import synth
from oracle import ssd_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _inputs(seed, B):
    pri = H.priors()
    gb, gc = synth.make_gt(seed, B)
    loc, conf = synth.make_head(seed, B, pri.shape[0])
    gx, gcl, off = synth.pack_gt(gb, gc)
    return pri, loc, conf, gb, gc, gx, gcl, off


@pytest.mark.parametrize("B", [3, 32])
def test_loss_host_matches_oracle_and_device_path(B):
    from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty, pinned_free
    pri, loc, conf, gb, gc, gx, gcl, off = _inputs(21, B)
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    hl[:] = loc
    hc[:] = conf
    gl, gcf = pinned_empty(loc.shape), pinned_empty(conf.shape)
    l1, l2 = ctx.loss_host(hl, hc, gx, gcl, off, gl, gcf)
    f1, f2 = ctx.loss_host(hl, hc, gx, gcl, off)           # forward only
    tb = [torch.from_numpy(b) for b in gb]
    tc = [torch.from_numpy(c) for c in gc]
    tl, tcf = torch.from_numpy(loc), torch.from_numpy(conf)
    ref = O.multibox_loss(tl, tcf, tb, tc, pri)
    assert abs(l1 - ref["loc_loss"].item()) <= 1e-5 * abs(ref["loc_loss"].item())
    assert abs(l2 - ref["conf_loss"].item()) <= 1e-5 * abs(ref["conf_loss"].item())
    assert (f1, f2) == (l1, l2)
    rgl, rgc = O.multibox_grads(tl, tcf, ref)
    assert torch.allclose(torch.from_numpy(gl.copy()), rgl, rtol=1e-5, atol=1e-9)
    got = torch.from_numpy(gcf.copy())
    same_rows = torch.equal(got.abs().sum(-1) != 0, rgc.abs().sum(-1) != 0)
    if same_rows:                                          # mined sets agree (no ulp-level boundary flip)
        assert torch.allclose(got, rgc, rtol=1e-4, atol=1e-8)
    # bit-identical to the step-by-step device path
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    head = MultiboxHead(pri, "cuda")
    out = head.loss(tl.cuda(), tcf.cuda(), PackedGT(tb, tc, head.dev), with_grads=True)
    assert torch.equal(out["grad_conf"].cpu(), got)
    assert torch.equal(out["grad_loc"].cpu(), torch.from_numpy(gl.copy()))
    assert abs(out["losses"][0].item() - l1) <= 1e-6 * abs(l1) and abs(out["losses"][1].item() - l2) <= 1e-6 * abs(l2)
    for a in (hl, hc, gl, gcf):
        pinned_free(a)
    ctx.close()


def test_loss_dev_single_call():
    from objectdetection_ssd_b200.ctx import SSDHeadContext
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    B = 16
    pri, loc, conf, gb, gc, gx, gcl, off = _inputs(22, B)
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tl, tcf, tgx, tgc, toff = d(loc), d(conf), d(gx), d(gcl), d(off)
    sums = torch.empty(2, dtype=torch.float64, device="cuda")
    losses = torch.empty(2, device="cuda")
    gl, gcf = torch.empty_like(tl), torch.empty_like(tcf)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ctx.loss_dev(tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]),
                     sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
    torch.cuda.synchronize()
    head = MultiboxHead(pri, "cuda")
    out = head.loss(tl, tcf, PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev),
                    with_grads=True)
    torch.cuda.synchronize()
    assert torch.equal(out["losses"], losses) and torch.equal(out["sums"], sums)
    assert torch.equal(out["grad_conf"], gcf) and torch.equal(out["grad_loc"], gl)
    ctx.close()


def test_detect_host_matches_device_path_and_survives_batch_changes():
    from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty
    from objectdetection_ssd_b200.head import MultiboxHead, detect
    pri = H.priors()
    P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=6)
    head = MultiboxHead(pri, "cuda")
    for B in (6, 2, 5):                                   # one context, changing batch sizes
        loc, conf = synth.make_head(80 + B, B, P, loc_scale=0.5, bg_bias=7.5)
        hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
        hl[:] = loc
        hc[:] = conf
        ob, op = pinned_empty((B, 200, 4)), pinned_empty((B, 200))
        oc, oi, on = pinned_empty((B, 200), np.int32), pinned_empty((B, 200), np.int32), pinned_empty((B,), np.int32)
        ctx.detect_host(hl, hc, ob, op, oc, oi, on, 0.01, 0.45)
        out = detect(head, torch.from_numpy(loc), torch.from_numpy(conf), 0.01, 0.45, 200)
        torch.cuda.synchronize()
        assert np.array_equal(on, out["cnt"].cpu().numpy())
        for b in range(B):
            k = int(on[b])
            assert np.array_equal(oi[b, :k], out["prior"][b, :k].cpu().numpy())
            assert np.array_equal(oc[b, :k], out["cls"][b, :k].cpu().numpy())
            assert np.array_equal(op[b, :k], out["prob"][b, :k].cpu().numpy())
            assert np.array_equal(ob[b, :k], out["boxes"][b, :k].cpu().numpy())
    ctx.close()


def test_host_paths_with_pageable_buffers_equal_pinned_ones():
    """Page-locked `loc` is read in place by the kernels (UVA alias); ordinary numpy arrays take the copy path.  Both
    must give the same losses, gradients and detections, bit for bit."""
    import numpy as np
    from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty, pinned_free
    B = 5
    pri, loc, conf, gb, gc, gx, gcl, off = _inputs(23, B)
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    hl[:] = loc
    hc[:] = conf
    g1l, g1c = np.empty_like(loc), np.empty_like(conf)
    g2l, g2c = np.empty_like(loc), np.empty_like(conf)
    # page-locked gradient buffers full of garbage: host threads must zero every element the GPU does not write
    p1l, p1c = pinned_empty(loc.shape), pinned_empty(conf.shape)
    p1l[:] = 7.0
    p1c[:] = -3.0
    s = ctx.loss_host(hl, hc, gx, gcl, off, p1l, p1c)                                   # sparse return path
    a = ctx.loss_host(hl, hc, gx, gcl, off, g1l, g1c)                                   # pinned inputs
    assert s == a and np.array_equal(p1l, g1l) and np.array_equal(p1c, g1c)
    pinned_free(p1l)
    pinned_free(p1c)
    b = ctx.loss_host(np.ascontiguousarray(loc), np.ascontiguousarray(conf), gx, gcl, off, g2l, g2c)   # pageable inputs
    assert a == b and np.array_equal(g1l, g2l) and np.array_equal(g1c, g2c)
    outs = []
    for l_, c_ in ((hl, hc), (np.ascontiguousarray(loc), np.ascontiguousarray(conf))):
        ob, op = np.empty((B, 200, 4), np.float32), np.empty((B, 200), np.float32)
        oc, oi, on = np.empty((B, 200), np.int32), np.empty((B, 200), np.int32), np.empty((B,), np.int32)
        ctx.detect_host(l_, c_, ob, op, oc, oi, on, 0.05, 0.45)
        outs.append((ob, op, oc, oi, on))
    for i in range(B):
        k = int(outs[0][4][i])
        assert k == int(outs[1][4][i])
        for x, y in zip(outs[0][:4], outs[1][:4]):
            assert np.array_equal(x[i, :k], y[i, :k])
    pinned_free(hl)
    pinned_free(hc)
    ctx.close()


@pytest.mark.parametrize("B,pinned_rows", [(5, True), (32, True), (32, False)])
def test_loss_host_sparse_rows_equal_the_dense_gradients(B, pinned_rows):
    """ssdhead_ctx_multibox_loss_host_sparse: the packed rows, scattered, are bit-identical to the dense gradient buffers
    of ssdhead_ctx_multibox_loss_host; positives come first and own the loc rows; a too small row_cap is reported
    through the counts (rows beyond it are dropped, nothing is written out of bounds)."""
    from objectdetection_ssd_b200.ctx import SSDHeadContext, SparseRows, pinned_empty
    pri, loc, conf, gb, gc, gx, gcl, off = _inputs(23, B)
    P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    hl[:] = loc
    hc[:] = conf
    gl, gcf = pinned_empty(loc.shape), pinned_empty(conf.shape)
    l1, l2 = ctx.loss_host(hl, hc, gx, gcl, off, gl, gcf)
    rows = SparseRows(B, cap=1024, pinned=pinned_rows)
    rows.idx[...] = -1
    s1, s2 = ctx.loss_host_sparse(hl, hc, gx, gcl, off, rows)
    assert (s1, s2) == (l1, l2)
    sgl, sgc = rows.scatter(P)
    assert np.array_equal(sgl, gl) and np.array_equal(sgc, gcf)
    for b in range(B):
        n, npos = rows.cnt[b]
        assert n == int((gcf[b] != 0).any(-1).sum()) and npos == int((gl[b] != 0).any(-1).sum())
        assert len(set(rows.idx[b, :n].tolist())) == n
        assert not pinned_rows or (rows.idx[b, n:] == -1).all()      # in-place rows: nothing is written past the count
    assert rows.nbytes_used() < 0.1 * (gl.nbytes + gcf.nbytes)
    # a cap below the largest image: counts still report the true number, the stored rows are a subset
    small = int(rows.cnt[:, 0].max()) - 3
    r2 = SparseRows(B, cap=small, pinned=pinned_rows)
    r2.idx[...] = -1
    ctx.loss_host_sparse(hl, hc, gx, gcl, off, r2)
    assert np.array_equal(r2.cnt, rows.cnt)
    b = int(rows.cnt[:, 0].argmax())
    assert set(r2.idx[b].tolist()) <= set(rows.idx[b, :rows.cnt[b, 0]].tolist())
    with pytest.raises(RuntimeError):
        r2.scatter(P)
    ctx.close()


def test_resident_gradient_tensors_equal_the_dense_step():
    """ssdhead_ctx_multibox_loss_dev_resident: the same gradient tensors from step to step, every step retracts the
    previous step's rows and writes its own - after every step the tensors must hold exactly what the dense step
    (zero background + rows, ssdhead_ctx_multibox_loss_dev) writes: same bits, same losses.  Covers changing batches,
    a smaller batch between larger ones, repeated inputs (the same rows written again), and the `fresh` hand-over
    after somebody else scribbled over the tensors."""
    from objectdetection_ssd_b200.ctx import SSDHeadContext
    from objectdetection_ssd_b200 import _lib
    maxB = 12
    pri = H.priors()
    P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=maxB)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    gl_res = torch.full((maxB, P, 4), 7.0, device="cuda")            # garbage: `fresh` must clean it
    gc_res = torch.full((maxB, P, 21), -3.0, device="cuda")
    sums_r = torch.empty(2, dtype=torch.float64, device="cuda")
    loss_r = torch.empty(2, device="cuda")
    sums_d = torch.empty(2, dtype=torch.float64, device="cuda")
    loss_d = torch.empty(2, device="cuda")
    plan = [(31, 12, True), (32, 12, False), (32, 12, False), (33, 5, False), (34, 12, False), (35, 12, True), (36, 1, False), (31, 12, False)]
    for step, (seed, B, fresh) in enumerate(plan):
        _, loc, conf, gb, gc, gx, gcl, off = _inputs(seed, B)
        tl, tcf, tgx, tgc, toff = d(loc), d(conf), d(gx), d(gcl), d(off)
        if fresh and step > 0:
            gc_res.fill_(1.0)                                        # somebody else wrote the tensors
            gl_res.fill_(2.0)
        ctx.loss_dev_resident(tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]),
                              sums_r.data_ptr(), loss_r.data_ptr(), gl_res.data_ptr(), gc_res.data_ptr(), st, fresh=fresh)
        gl_d, gc_d = torch.empty_like(tl), torch.empty_like(tcf)
        ctx.loss_dev(tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]),
                     sums_d.data_ptr(), loss_d.data_ptr(), gl_d.data_ptr(), gc_d.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.equal(loss_r, loss_d) and torch.equal(sums_r, sums_d), f"step {step}: losses"
        assert torch.equal(gc_res[:B], gc_d), f"step {step}: grad_conf"
        assert torch.equal(gl_res[:B], gl_d), f"step {step}: grad_loc"
        assert int((gc_d.abs().sum(-1) != 0).sum()) > 0
    # other tensors without `fresh`: refused, nothing launched
    other = torch.zeros_like(gc_res)
    rc = ctx.lib.ssdhead_ctx_multibox_loss_dev_resident(ctx._h, tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(),
                                                        toff.data_ptr(), B, int(off[-1]), 3, 0.5, sums_r.data_ptr(), loss_r.data_ptr(),
                                                        gl_res.data_ptr(), other.data_ptr(), 0, st)
    assert rc == -5
    ctx.close()
    # a batch larger than the one the tensors were handed over with: refused as well (only that many images are clean)
    ctx = SSDHeadContext(pri.numpy(), max_batch=maxB)
    _, loc, conf, gb, gc, gx, gcl, off = _inputs(40, 4)
    tl, tcf, tgx, tgc, toff = d(loc), d(conf), d(gx), d(gcl), d(off)
    ctx.loss_dev_resident(tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), 4, int(off[-1]),
                          sums_r.data_ptr(), loss_r.data_ptr(), gl_res.data_ptr(), gc_res.data_ptr(), st, fresh=True)
    _, loc, conf, gb, gc, gx, gcl, off = _inputs(41, 8)
    tl, tcf, tgx, tgc, toff = d(loc), d(conf), d(gx), d(gcl), d(off)
    rc = ctx.lib.ssdhead_ctx_multibox_loss_dev_resident(ctx._h, tl.data_ptr(), tcf.data_ptr(), tgx.data_ptr(), tgc.data_ptr(),
                                                        toff.data_ptr(), 8, int(off[-1]), 3, 0.5, sums_r.data_ptr(), loss_r.data_ptr(),
                                                        gl_res.data_ptr(), gc_res.data_ptr(), 0, st)
    assert rc == -5
    torch.cuda.synchronize()
    ctx.close()
