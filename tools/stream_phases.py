"""Developer tool: per-CTA statistics of the detect stream kernel (items, clocks the first consumer warp waits for data and
works, clocks the producer waits for a free stage and spends on bookkeeping, sampling prologue, start/end on the global timer).

Needs a developer build:  SSDHEAD_NVCC_EXTRA=-DSSDHEAD_PHASE_TIMES python -m objectdetection_ssd_b200.build --force
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.head import MultiboxHead, detect

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
G = int(sys.argv[2]) if len(sys.argv) > 2 else 592
pri = PR.make_priors()
head = MultiboxHead(pri, "cuda")
loc, conf = synth.make_head(3, B, pri.shape[0], loc_scale=0.5, bg_bias=6.0)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
lib = ctypes.CDLL(_lib.LIB_PATH)
G = min(B * 35, G)
clk = 1965.0
q = lambda v: "min %.2f med %.2f max %.2f" % (np.min(v), np.median(v), np.max(v))
for it in range(3):
    out = detect(head, l, c, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (8 * G))()
    lib.ssdhead_debug_stream(buf, G)
    a = np.array(list(buf), dtype=np.int64).reshape(G, 8).astype(np.float64)
    a = a[a[:, 0] > 0]
    t0 = a[:, 5].min()
    n = a[:, 0]
    print(f"B={B} run {it}: {len(a)} CTAs; producers start {q((a[:, 5] - t0) / 1e3)} us; consumers end {q((a[:, 6] - t0) / 1e3)} us; items/CTA {q(n)}")
    print("   per item (us): consumer waits for data", q(a[:, 1] / n / clk), "| works", q(a[:, 2] / n / clk),
          "| producer waits for a stage", q(a[:, 3] / n / clk), "| bookkeeping", q(a[:, 4] / n / clk), "| sampling prologue (us)", q(a[:, 7] / clk))
