#!/usr/bin/env python
"""bench.py - SSD multibox head path on B200: images/sec and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload train|detect] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic head outputs:
  train  (default)  match + hard-negative-mined CE + L1 loss, forward AND gradients
                    (BASELINE.json configs[1]/[3]; default batch 256 per GPU = north_star's target size)
  detect            decode + conf 0.01 threshold + per-class NMS + global top-200 (configs[2], batch 64)
Under torchrun (N > 1) every rank owns `--batch` images (weak scaling); the batch-global positive
count and the loss sums are all-reduced over NCCL exactly as a sharded ssd() call does.
Rank 0 prints ONE JSON line.  `--impl reference` times the CPU restatement of the reference
(oracle/, torch-CPU ops, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

ALGO_BYTES = {  # SURVEY.md 8(d), fp32, P = 8732
    "train": 1_746_400,      # loc + conf in, dense grad_loc + grad_conf out
    "train_fwd": 873_200,
    "detect": 878_800,
}
# algorithmic bytes per image of the dominant kernel alone (DESIGN.md "Kernels"): ce_stream_kernel reads conf
# (8732*84) and writes CE (8732*4), one class byte per prior and the dense gradient background (8732*100)
CE_STREAM_BYTES = 8732 * (84 + 4 + 1 + 100)
L2_BYTES = 126 * 1024 * 1024


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._timed = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self._timed.is_set(), mhz))
                if self._timed.is_set():
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def timed(self, on: bool):
        (self._timed.set if on else self._timed.clear)()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        timed = [m for t, m in self.samples if t] or [m for _, m in self.samples]
        return {"sm_mhz": float(np.median(timed)) if timed else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples_in_timed_region": len([1 for t, _ in self.samples if t])}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------- reference arm
def reference_inputs(workload, B, P):
    from objectdetection_ssd_b200 import synth
    if workload == "train":
        gb, gc = synth.make_gt(1, B)
        loc, conf = synth.make_head(1, B, P)
        return (torch.from_numpy(loc), torch.from_numpy(conf), [torch.from_numpy(b) for b in gb],
                [torch.from_numpy(c) for c in gc])
    loc, conf = synth.make_head(3, B, P, loc_scale=0.5, bg_bias=6.0)
    return torch.from_numpy(loc), torch.from_numpy(conf)


def reference_step(workload, data, pri, pxy):
    """One pass of the reference's CPU algorithm (oracle port, torch-CPU ops) over the sample."""
    from oracle import ssd_oracle as O
    if workload == "train":
        loc, conf, tb, tc = data
        l = loc.clone().requires_grad_(True)
        c = conf.clone().requires_grad_(True)
        a, b = O.ssd_reference_style((l, c), tc, tb, pri, pxy)
        (a + b).backward()
        return a.item() + b.item()
    loc, conf = data
    n = 0
    for i in range(loc.shape[0]):
        n += O.detect_image(loc[i], conf[i], pri, 0.01, 0.45, 200)[0].shape[0]
    return n


def cpu_baseline(workload, budget_s=20.0):
    """Bounded CPU sample of the same workload (SURVEY.md 8(d) 'CPU baseline')."""
    from oracle import ssd_oracle as O
    pri = O.make_priors()
    pxy = O.cxcywh_to_xyxy(pri)
    B = 32 if workload == "train" else 1
    data = reference_inputs(workload, B, pri.shape[0])
    t0 = time.perf_counter()
    reference_step(workload, data, pri, pxy)            # warm-up
    first = time.perf_counter() - t0
    best = first
    reps = 0
    while reps < 3 and (time.perf_counter() - t0) + best < budget_s:
        t = time.perf_counter()
        reference_step(workload, data, pri, pxy)
        best = min(best, time.perf_counter() - t)
        reps += 1
    what = ("oracle.ssd_reference_style fwd+bwd, batch 32, 1-10 gt/image" if workload == "train"
            else "oracle.detect_image, 1 image, bg bias +6, min_score 0.01")
    return {"value": B / best, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{what}; best of {reps + 1} runs, {best * 1e3:.1f} ms/pass"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import ssd_oracle as O
    pri = O.make_priors()
    pxy = O.cxcywh_to_xyxy(pri)
    B = min(args.batch, 32) if args.workload == "train" else min(args.batch, 2)
    data = reference_inputs(args.workload, B, pri.shape[0])
    for _ in range(min(args.warmup, 1)):
        reference_step(args.workload, data, pri, pxy)
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(args.workload, data, pri, pxy)
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    sample = (f"batch {B} of the {args.batch}-image workload per step (bounded sample), {steps} timed steps"
              f" (requested {args.steps}), torch {torch.__version__} CPU ops")
    line = {
        "impl": "reference", "metric": "images/sec, SSD300 " + ("match + multibox loss fwd+bwd" if args.workload == "train"
                                                                else "decode + NMS (conf 0.01, top-200)"),
        "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    if args.workload == "train":
        return {"workload": "SSD300-VGG16 VOC head, 8732 priors, 21 classes: match + hard-negative-mined CE + L1 loss, "
                            "forward and gradients (north_star target size: batch 256 per GPU)",
                "batch_per_gpu": args.batch, "global_batch": args.batch * world, "num_priors": 8732, "num_classes": 21,
                "gt_per_image": "U{1..10}", "parallelism": (f"image-sharded x{world}; Npos and the loss sums cross GPUs "
                                + ("through NVLink peer-memory stores inside the mining kernel (--collective peer)"
                                   if getattr(args, "collective", "peer") == "peer" else "through two NCCL all-reduces (--collective nccl)")
                                if world > 1 else "single GPU"),
                "l2": "inputs rotate through >= 2 x 126 MB of distinct device buffers (larger than L2)"}
    return {"workload": "SSD300 inference post-processing: decode + conf 0.01 threshold + per-class NMS (iou 0.45) + "
                        "global top-200; background-logit bias +6 (~1.1k candidates/class)",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "num_priors": 8732, "num_classes": 21,
            "parallelism": f"image-sharded x{world}, no collective",
            "l2": "inputs rotate through >= 2 x 126 MB of distinct device buffers (larger than L2)"}


# ----------------------------------------------------------------------------------------------- our arm
def max_over_ranks(ms, world, dev):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


class _DevPtrView:
    """Zero-copy torch view of device memory owned by the library (for the NCCL all-reduce)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def time_train(args, rank, world, dev, sampler):
    from objectdetection_ssd_b200 import _lib, synth, priors as PR
    from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty
    lib = _lib.load()
    B = args.batch
    pri = PR.make_priors()
    P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=B, device=dev.index)
    gb, gc = synth.make_gt(1 + rank, B)
    gx, gcl, off = synth.pack_gt(gb, gc)
    loc, conf = synth.make_head(1 + rank, B, P)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tgx, tgc, toff = d(gx), d(gcl), d(off)
    per_set = B * P * 25 * 4
    nset = max(2, -(-2 * L2_BYTES // per_set))
    sets = [(d(loc) + 0.001 * i, d(conf) + 0.001 * i) for i in range(nset)]
    sums = torch.empty(2, dtype=torch.float64, device=dev)
    losses = torch.empty(2, device=dev)
    gl, gcf = torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    sumG = int(off[-1])
    npos_norm = torch.zeros(1, dtype=torch.int32, device=dev)
    collective = "none"
    if world > 1:
        import torch.distributed as dist
        collective = args.collective
        if collective == "peer" and B > 2 * torch.cuda.get_device_properties(dev).multi_processor_count:
            collective = "nccl"          # the in-kernel exchange needs one co-resident CTA per image (B <= 2 x SMs)
        if collective == "peer":
            # one-off: exchange the CUDA IPC handles of the ranks' exchange buffers; afterwards the step is the same
            # two kernels as on one GPU, the mining kernel trading Npos and the loss sums with its peers over NVLink
            handles = [None] * world
            dist.all_gather_object(handles, ctx.xchg_export())
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                ctx.xchg_import(handles, rank)
            except RuntimeError as e:                      # no peer access between these GPUs: every rank falls back
                print(f"bench.py: rank {rank}: peer-memory import failed ({e}); using --collective nccl", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                ctx.close()
                ctx = SSDHeadContext(pri.numpy(), max_batch=B, device=dev.index)   # a context without the exchange
                collective = "nccl"
            dist.barrier()
        args.collective = collective      # what actually ran (reported in config.parallelism)

    def step(i):
        l, c = sets[i % nset]
        if world == 1 or collective == "peer":
            ctx.loss_dev(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, sumG,
                         sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
        else:
            p = ctx.loss_begin(c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, sumG,
                               gl.data_ptr(), gcf.data_ptr(), st)
            npos_norm.copy_(torch.as_tensor(_DevPtrView(p, 1, "<i4"), device=dev))
            dist.all_reduce(npos_norm)
            ctx.loss_end(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B,
                         npos_norm.data_ptr(), sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
            dist.all_reduce(sums)
            ctx.finish_loss(sums.data_ptr(), npos_norm.data_ptr(), losses.data_ptr(), st)

    for i in range(args.warmup):
        step(i)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    sampler.timed(True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier(world)
    sampler.timed(False)
    launches = _lib.launch_count() - n0
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    loss_val = losses.tolist()
    if collective == "peer" and ctx.xchg_error():
        raise SystemExit("bench.py: a wait for a peer rank expired inside the sharded step (results invalid)")

    # ---- the dominant kernel alone, CUDA events on its launch stream: the streaming CE kernel with the fused natural
    # match (ssdhead_ce_match_stream with run_finalizer = 0, on a scratch match workspace) ----
    ws_bytes = int(lib.ssdhead_workspace_bytes(_lib.WS_LOSS, B, P, 21, 0))
    ws = torch.zeros(ws_bytes + 256, dtype=torch.uint8, device=dev)
    wm = torch.zeros(int(lib.ssdhead_workspace_bytes(_lib.WS_MATCH, B, P, 21, sumG)) + 256, dtype=torch.uint8, device=dev)
    pri_xyxy = PR.cxcywh_to_xyxy_host(pri).to(dev)
    cls_u8 = torch.empty(B, P, dtype=torch.uint8, device=dev)
    bestp = torch.empty(max(sumG, 1), dtype=torch.int32, device=dev)
    npos_k = torch.empty(B + 1, dtype=torch.int32, device=dev)

    def kern(i):
        return lib.ssdhead_ce_match_stream(sets[i % nset][1].data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(),
                                           pri_xyxy.data_ptr(), B, P, 21, sumG, 0.5, None, gl.data_ptr(), gcf.data_ptr(),
                                           cls_u8.data_ptr(), bestp.data_ptr(), npos_k.data_ptr(),
                                           ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), 0, st)

    kn = max(10, min(args.steps, 200))
    for i in range(3):
        _lib.check(kern(i), "ssdhead_ce_match_stream")
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for i in range(kn):
        kern(i)
    k1.record(stream)
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / kn

    # ---- end to end through the host-buffer C ABI (pinned host tensors in, losses + gradients out) ----
    hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    hl[:] = loc
    hc[:] = conf
    hgl, hgc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    for _ in range(2):
        ctx.loss_host(hl, hc, gx, gcl, off, hgl, hgc)
    en = max(3, min(args.steps, 20))
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(en):
        ctx.loss_host(hl, hc, gx, gcl, off, hgl, hgc)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / en, world, dev)
    # conf and gt are copied; loc is read in place from the page-locked host buffer, positive rows only (counted as 128
    # rows of 32-byte sectors per image, an upper bound at 1-10 gts per image)
    h2d = conf.nbytes + gx.nbytes + gcl.nbytes + off.nbytes + B * 128 * 32
    # the dense zero background of the gradients never crosses PCIe when the caller's buffers are page-locked: host threads
    # zero them while the inputs stream in and the mining kernel stores its rows straight into them.
    # Bytes counted from what arrived: the non-zero rows of the two host gradient tensors + the two losses.
    if os.environ.get("SSDHEAD_E2E_SPARSE", "1") != "0":
        d2h = int((hgc != 0).any(-1).sum()) * 84 + int((hgl != 0).any(-1).sum()) * 16 + 8
    else:
        d2h = loc.nbytes + conf.nbytes + 8
    ctx.close()
    return dict(ms_total=ms, launches=launches, kern_ms=kern_ms, e2e_ms=e2e_ms, h2d=h2d, d2h=d2h, losses=loss_val,
                kernel="ce_stream_kernel<21,true,true>", kernel_bytes=CE_STREAM_BYTES * B, algo=ALGO_BYTES["train"],
                e2e_steps=en, collective=collective)


def time_detect(args, rank, world, dev, sampler, steps=None, warmup=None):
    from objectdetection_ssd_b200 import _lib, synth, priors as PR
    from objectdetection_ssd_b200.head import MultiboxHead, detect
    from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty
    steps = steps or args.steps
    warmup = warmup if warmup is not None else args.warmup
    B = args.batch if args.workload == "detect" else 64
    pri = PR.make_priors()
    P = pri.shape[0]
    head = MultiboxHead(pri, dev)
    loc, conf = synth.make_head(3 + rank, B, P, loc_scale=0.5, bg_bias=6.0)
    per_set = B * P * 25 * 4
    nset = max(2, -(-2 * L2_BYTES // per_set))
    sets = [(torch.from_numpy(loc).to(dev) + 1e-4 * i, torch.from_numpy(conf).to(dev)) for i in range(nset)]
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    lib = _lib.load()
    top_k = 200
    out = dict(boxes=torch.empty(B, top_k, 4, device=dev), prob=torch.empty(B, top_k, device=dev),
               cls=torch.empty(B, top_k, dtype=torch.int32, device=dev),
               prior=torch.empty(B, top_k, dtype=torch.int32, device=dev),
               cnt=torch.empty(B, dtype=torch.int32, device=dev))
    ws = torch.zeros(int(lib.ssdhead_workspace_bytes(_lib.WS_DETECT, B, P, 21, 0)) + 256, dtype=torch.uint8, device=dev)
    pri_dev = head.pri_cxcywh

    def step(i):
        # the C ABI directly (what objectdetection_ssd_b200.head.detect calls after allocating its outputs)
        l, c = sets[i % nset]
        return lib.ssdhead_detect(l.data_ptr(), c.data_ptr(), pri_dev.data_ptr(), B, P, 21, 0.01, 0.45, top_k, None, 0,
                                  out["boxes"].data_ptr(), out["prob"].data_ptr(), out["cls"].data_ptr(),
                                  out["prior"].data_ptr(), out["cnt"].data_ptr(), ws.data_ptr(), ws.numel(), st)

    for i in range(max(warmup, 1)):
        _lib.check(step(i), "ssdhead_detect")
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    sampler.timed(True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    e1.record(stream)
    barrier(world)
    sampler.timed(False)
    launches = _lib.launch_count() - n0
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    # end to end: host buffers in, detections out
    ctx = SSDHeadContext(pri.numpy(), max_batch=B, device=dev.index)
    hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape)
    hl[:] = loc
    hc[:] = conf
    ob, op = pinned_empty((B, 200, 4)), pinned_empty((B, 200))
    oc, oi, on = pinned_empty((B, 200), np.int32), pinned_empty((B, 200), np.int32), pinned_empty((B,), np.int32)
    for _ in range(2):
        ctx.detect_host(hl, hc, ob, op, oc, oi, on, 0.01, 0.45)
    en = max(3, min(steps, 20))
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(en):
        ctx.detect_host(hl, hc, ob, op, oc, oi, on, 0.01, 0.45)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / en, world, dev)
    ctx.close()
    # conf is copied; of loc only the rows of the candidates the sweep visits cross PCIe (the kernel reads the page-locked
    # host buffer in place): counted as 512 rows of 32-byte sectors per image, an upper bound at this workload
    return dict(ms_total=ms, launches=launches, kern_ms=None, e2e_ms=e2e_ms, h2d=conf.nbytes + B * 512 * 32,
                d2h=ob.nbytes + op.nbytes + oc.nbytes + oi.nbytes + on.nbytes, B=B, steps=steps,
                kernel="detect_score_kernel + detect_nms_kernel (whole step)", algo=ALGO_BYTES["detect"], e2e_steps=en,
                detections=int(out["cnt"].clamp(min=0).sum()))


def time_train_levels(rank, dev, B=256, steps=200):
    """Training-head step on the six per-level tensors (NHWC conv outputs viewed as rows) instead of the concatenated
    [B,8732,*] pair - SURVEY.md 8(f) #3 - next to what building that pair costs (the torch.cat of Model.py:234-235)."""
    import ctypes
    from objectdetection_ssd_b200 import _lib, synth, priors as PR
    from objectdetection_ssd_b200.head import MultiboxHead, PackedGT
    lib = _lib.load()
    pri = PR.make_priors()
    P = pri.shape[0]
    head = MultiboxHead(pri, dev)
    gb, gc = synth.make_gt(1 + rank, B)
    gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
    loc, conf = synth.make_head(1 + rank, B, P)
    counts = (5776, 2166, 600, 150, 36, 4)
    nset = max(2, -(-2 * L2_BYTES // (B * P * 25 * 4)))
    sets, structs = [], []
    for i in range(nset):
        l, c = torch.from_numpy(loc).to(dev) + 0.001 * i, torch.from_numpy(conf).to(dev) + 0.001 * i
        ls, cs, s0 = [], [], 0
        for n in counts:
            ls.append(l[:, s0:s0 + n].contiguous())
            cs.append(c[:, s0:s0 + n].contiguous())
            s0 += n
        sets.append((ls, cs))
    gls = [torch.empty_like(t) for t in sets[0][0]]
    gcs = [torch.empty_like(t) for t in sets[0][1]]
    for ls, cs in sets:
        st_ = _lib.Levels()
        st_.num_levels = len(counts)
        for i in range(len(counts)):
            st_.count[i] = counts[i]
            st_.conf[i], st_.loc[i] = cs[i].data_ptr(), ls[i].data_ptr()
            st_.grad_conf[i], st_.grad_loc[i] = gcs[i].data_ptr(), gls[i].data_ptr()
        structs.append(st_)
    sums = torch.empty(2, dtype=torch.float64, device=dev)
    losses = torch.empty(2, device=dev)
    m = head._match_outputs(gt, False)
    ws = head._workspace(_lib.WS_LOSS, B, 0)
    wm = head._workspace(_lib.WS_MATCH, B, gt.sumG)
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream

    def step(i):
        return lib.ssdhead_multibox_step_levels(
            ctypes.addressof(structs[i % nset]), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
            head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, gt.sumG, 3, 0.5, sums.data_ptr(), losses.data_ptr(),
            m["cls_u8"].data_ptr(), m["best_prior"].data_ptr(), m["npos"].data_ptr(),
            ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), st)

    def cat(i):
        ls, cs = sets[i % nset]
        return torch.cat(ls, 1), torch.cat(cs, 1)

    def timeit(f):
        for i in range(5):
            f(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            f(i)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    _lib.check(step(0), "ssdhead_multibox_step_levels")
    return dict(ms_step=timeit(step), ms_cat=timeit(cat), losses=losses.tolist())


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the SSD head path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = peaks()
    sampler = ClockSampler(local)
    sampler.start()
    B = args.batch
    if args.workload == "train":
        r = time_train(args, rank, world, dev, sampler)
    else:
        r = time_detect(args, rank, world, dev, sampler)
    clocks = sampler.stop()
    ms_step = r["ms_total"] / args.steps
    value = B * world / (ms_step * 1e-3)
    e2e_val = B * world / (r["e2e_ms"] * 1e-3)
    step_gbs = r["algo"] * B / (ms_step * 1e-3) / 1e9
    line = {
        "metric": "images/sec, SSD300 " + ("match + multibox loss fwd+bwd" if args.workload == "train"
                                           else "decode + NMS (conf 0.01, top-200)"),
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(r["h2d"]), "d2h_bytes_per_step": int(r["d2h"]),
                "ms_per_step": r["e2e_ms"], "steps": r["e2e_steps"],
                "api": "ssdhead_ctx_multibox_loss_host" if args.workload == "train" else "ssdhead_ctx_detect_host",
                "note": ("pinned host buffers in; losses + dense gradient tensors in the caller's host buffers: conf is copied, "
                         "loc is read in place (positive rows only), the gradients' zero background is written by host threads "
                         "while the inputs stream in and the mining kernel stores its ~4 Npos rows straight into the host buffers"
                         if args.workload == "train" else
                         "pinned host buffers in, detections back to host: conf is copied, loc is read in place (visited candidates only)")
                        + "; per-rank call" + (", local normalisation" if world > 1 and args.workload == "train" else "")},
        "gpu_launches": int(r["launches"]),
        "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                          "algorithmic_bytes_per_image": r["algo"], "peak_source": peak_src,
                          "note": "whole step (all kernels of the step) against SURVEY.md 8(d) bytes/image"},
    }
    if r["kern_ms"] is not None:
        ach = r["kernel_bytes"] / (r["kern_ms"] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            ent = tj.get(f"ce_stream_b{B}")
            if ent:
                traffic = ent["dram_bytes_per_launch"]
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": traffic, "kernel": r["kernel"], "kernel_us": r["kern_ms"] * 1e3,
                            "algorithmic_bytes_per_launch": r["kernel_bytes"], "peak_source": peak_src}
    else:
        line["roofline"] = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                            "traffic": None, "kernel": r["kernel"],
                            "note": "whole step: the score kernel is issue-bound at ~22k candidates/image, the sweep kernel latency-bound (one CTA per image); see DESIGN.md",
                            "peak_source": peak_src}
    if "losses" in r:
        line["loss"] = r["losses"]
    if rank == 0 and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.workload)
        if args.workload == "train" and not args.no_others:
            # the other single-GPU configurations of BASELINE.json, short runs, for the record
            others = []
            for b2 in (32,):
                a2 = argparse.Namespace(**vars(args))
                a2.batch, a2.steps, a2.warmup = b2, 200, 10
                s2 = ClockSampler(local)
                r2 = time_train(a2, rank, world, dev, s2)
                m2 = r2["ms_total"] / a2.steps
                others.append({"workload": f"train head, batch {b2} (configs[1])", "images_per_s": b2 / (m2 * 1e-3),
                               "ms_per_step": m2, "step_roofline_frac": ALGO_BYTES['train'] * b2 / (m2 * 1e-3) / 1e9 / peak,
                               "e2e_images_per_s": b2 / (r2["e2e_ms"] * 1e-3)})
            a3 = argparse.Namespace(**vars(args))
            a3.workload, a3.batch = "detect", 64
            r3 = time_detect(a3, rank, world, dev, ClockSampler(local), steps=20, warmup=3)
            m3 = r3["ms_total"] / 20
            others.append({"workload": "detect, batch 64, bias +6 (configs[2])", "images_per_s": 64 / (m3 * 1e-3),
                           "ms_per_step": m3, "step_roofline_frac": ALGO_BYTES['detect'] * 64 / (m3 * 1e-3) / 1e9 / peak,
                           "e2e_images_per_s": 64 / (r3["e2e_ms"] * 1e-3)})
            r4 = time_train_levels(rank, dev)
            others.append({"workload": "train head, batch 256, from the 6 per-level tensors (ssdhead_multibox_step_levels, "
                                       "SURVEY 8(f) #3; no concatenated tensor in either direction)",
                           "images_per_s": 256 / (r4["ms_step"] * 1e-3), "ms_per_step": r4["ms_step"],
                           "step_roofline_frac": ALGO_BYTES['train'] * 256 / (r4["ms_step"] * 1e-3) / 1e9 / peak,
                           "torch_cat_of_the_levels_ms": r4["ms_cat"],
                           "note": "the concatenated layout pays value's step PLUS torch_cat_of_the_levels_ms (Model.py:234-235) "
                                   "and the same again in backward"})
            line["others"] = others
    elif rank == 0:
        line["cpu_baseline"] = None
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "detect"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 256 train, 64 detect)")
    ap.add_argument("--no-others", action="store_true", help="skip the short secondary-configuration runs")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how Npos and the loss sums cross GPUs - 'peer' = stores into peer memory over NVLink "
                         "from inside the mining kernel (2 kernels/step, no NCCL call), 'nccl' = two NCCL all-reduces "
                         "around the mining kernel (3 kernels/step)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 256 if args.workload == "train" else 64
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
