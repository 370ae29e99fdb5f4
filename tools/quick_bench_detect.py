"""Developer timing loop for the detect path (device-resident inputs, CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.head import MultiboxHead, detect

def run(B, bias, steps=10, warm=3):
    pri = PR.make_priors(); P = pri.shape[0]
    head = MultiboxHead(pri, "cuda")
    loc, conf = synth.make_head(3, B, P, loc_scale=0.5, bg_bias=bias)
    nset = max(1, int(300e6 // (B * P * 25 * 4)) + 1)
    sets = [(torch.from_numpy(loc).cuda() + 0.001 * i, torch.from_numpy(conf).cuda()) for i in range(nset)]
    for i in range(warm): out = detect(head, *sets[i % nset], 0.01, 0.45, 200)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): out = detect(head, *sets[i % nset], 0.01, 0.45, 200)
    e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1) / steps
    print(json.dumps(dict(B=B, bias=bias, step_us=round(tt * 1e3, 1), img_per_s=round(B / (tt * 1e-3)),
                          frac_of_6538=round(B * 878800 / (tt * 1e-3) / 1e9 / 6538.6, 4), cnt=out["cnt"][:4].tolist())))

if __name__ == "__main__":
    for bias in (8.0, 6.0, 4.0):
        run(64, bias)
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), 6.0)
