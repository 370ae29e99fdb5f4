"""Developer timing loop (not the contract bench): device-resident inputs, CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT

def run(B, steps=20, warm=5, grads=True):
    pri = PR.make_priors()
    P = pri.shape[0]
    head = MultiboxHead(pri, "cuda")
    gb, gc = synth.make_gt(1, B)
    loc, conf = synth.make_head(1, B, P)
    gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
    nset = max(1, int(400e6 // (B * P * 25 * 4)) + 1)
    sets = [(torch.from_numpy(loc).cuda() + i, torch.from_numpy(conf).cuda() + 0.01 * i) for i in range(nset)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for i in range(warm):
        head.loss(*sets[i % nset], gt, with_grads=grads)
    torch.cuda.synchronize()
    tt = 0.0
    for i in range(steps):
        l, c = sets[i % nset]
        ev[0].record()
        head.loss(l, c, gt, with_grads=grads)
        ev[2].record()
        torch.cuda.synchronize()
        tt += ev[0].elapsed_time(ev[2])
    tt /= steps
    bytes_img = 1746400 if grads else 873200
    print(json.dumps(dict(B=B, grads=grads, step_us=round(tt * 1e3, 1), img_per_s=round(B / (tt * 1e-3)),
                          frac_of_6538=round(B * bytes_img / (tt * 1e-3) / 1e9 / 6538.6, 3))))

if __name__ == "__main__":
    for B in (32, 64, 256):
        run(B, grads=True)
    run(256, grads=False)
