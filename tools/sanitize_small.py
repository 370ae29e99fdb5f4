"""Small end-to-end calls for compute-sanitizer (memcheck / racecheck): detect (both twins), loss step, per-level step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT, detect, detect_from_scores, detect_levels

pri = PR.make_priors(); P = pri.shape[0]
head = MultiboxHead(pri, "cuda")
B = 2
for bias in (6.0, 2.0):
    loc, conf = synth.make_head(3, B, P, loc_scale=0.5, bg_bias=bias)
    l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
    out = detect(head, l, c, 0.01, 0.45, 200)
    out2 = detect_from_scores(head, l, F.softmax(c, 2), 0.02, 0.45, 50)
    counts = (5776, 2166, 600, 150, 36, 4)
    ls, cs, s = [], [], 0
    for n in counts:
        ls.append(l[:, s:s + n].contiguous()); cs.append(c[:, s:s + n].contiguous()); s += n
    out3 = detect_levels(head, ls, cs, 0.01, 0.45, 200)
    torch.cuda.synchronize()
    print("detect", bias, out["cnt"].tolist(), out2["cnt"].tolist(), out3["cnt"].tolist())
gb, gc = synth.make_gt(1, B)
gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
loc, conf = synth.make_head(1, B, P)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
o = head.loss(l, c, gt, with_grads=True)
ls, cs, s = [], [], 0
for n in counts:
    ls.append(l[:, s:s + n].contiguous()); cs.append(c[:, s:s + n].contiguous()); s += n
o2 = head.loss_levels(ls, cs, gt, with_grads=True)
torch.cuda.synchronize()
print("loss", o["losses"].tolist(), o2["losses"].tolist())
