"""Single-shot driver for ncu: a few match+loss steps at a given batch (device-resident inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pri = PR.make_priors()
head = MultiboxHead(pri, "cuda")
gb, gc = synth.make_gt(1, B)
loc, conf = synth.make_head(1, B, pri.shape[0])
gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
for _ in range(steps):
    out = head.loss(l, c, gt, with_grads=True)
torch.cuda.synchronize()
print("ok", out["losses"].tolist())
