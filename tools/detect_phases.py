"""Developer tool: SM-clock phase times of the first image's sweep CTA (detect_nms_kernel).

Needs a developer build:  SSDHEAD_NVCC_EXTRA=-DSSDHEAD_PHASE_TIMES python -m objectdetection_ssd_b200.build --force
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.head import MultiboxHead, detect

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bias = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
pri = PR.make_priors()
head = MultiboxHead(pri, "cuda")
loc, conf = synth.make_head(3, B, pri.shape[0], loc_scale=0.5, bg_bias=bias)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
lib = _lib.load()
names = ["wait->start", "coarse scan", "filter", "sort", "decode", "sweep", "emit"]
for it in range(4):
    out = detect(head, l, c, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    rc = lib.ssdhead_debug_phases(buf)
    full = list(buf)
    ph = full[:7]
    print(it, " ".join(f"{names[i + 1]}={(ph[i + 1] - ph[i]) / 1965.0:.2f}us" for i in range(6)), f"total={(ph[6] - ph[0]) / 1965.0:.2f}us",
          "| sweep: partition=%.2f classes=%.2f compact=%.2f" % tuple((full[i + 1] - full[i]) / 1965.0 for i in (8, 9, 10)))
