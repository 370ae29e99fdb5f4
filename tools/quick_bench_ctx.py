"""Developer timing loop through the one-call context (device-resident inputs, CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.ctx import SSDHeadContext

def run(B, steps=30, warm=5, grads=True):
    pri = PR.make_priors(); P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    gb, gc = synth.make_gt(1, B); gx, gcl, off = synth.pack_gt(gb, gc)
    loc, conf = synth.make_head(1, B, P)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tgx, tgc, toff = d(gx), d(gcl), d(off)
    nset = max(1, int(400e6 // (B * P * 25 * 4)) + 1)
    sets = [(d(loc) + i, d(conf) + 0.01 * i) for i in range(nset)]
    sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
    gl = torch.empty_like(sets[0][0]) if grads else None; gcf = torch.empty_like(sets[0][1]) if grads else None
    st = torch.cuda.current_stream().cuda_stream
    def step(i):
        l, c = sets[i % nset]
        ctx.loss_dev(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]),
                     sums.data_ptr(), losses.data_ptr(), gl.data_ptr() if grads else None, gcf.data_ptr() if grads else None, st)
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): step(i)
    e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1) / steps
    bytes_img = 1746400 if grads else 873200
    print(json.dumps(dict(B=B, grads=grads, step_us=round(tt * 1e3, 1), img_per_s=round(B / (tt * 1e-3)),
                          frac_of_6538=round(B * bytes_img / (tt * 1e-3) / 1e9 / 6538.6, 3))))
    ctx.close()

if __name__ == "__main__":
    for B in (32, 64, 128, 256):
        run(B)
    run(256, grads=False)
