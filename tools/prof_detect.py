"""Single-shot driver for ncu: a few detect steps (B=64, bias +6)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.head import MultiboxHead, detect
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bias = float(sys.argv[2]) if len(sys.argv) > 2 else 6.0
pri = PR.make_priors()
head = MultiboxHead(pri, "cuda")
loc, conf = synth.make_head(3, B, pri.shape[0], loc_scale=0.5, bg_bias=bias)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
for _ in range(3):
    out = detect(head, l, c, 0.01, 0.45, 200)
torch.cuda.synchronize()
print("ok", out["cnt"][:4].tolist())
