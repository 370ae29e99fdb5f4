"""Developer timing of the whole training step (ctx.loss_dev), launches back to back like bench.py: several repeats,
min / median per step, plus an exact checksum of the outputs so that library variants can be compared bit for bit.

usage: quick_bench_step.py [B ...]     (default 256 32)
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.ctx import SSDHeadContext


def run(B, steps=500, reps=6):
    pri = PR.make_priors(); P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    gb, gc = synth.make_gt(1, B); gx, gcl, off = synth.pack_gt(gb, gc)
    loc, conf = synth.make_head(1, B, P)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tgx, tgc, toff = d(gx), d(gcl), d(off)
    nset = max(2, int(2 * 126e6 // (B * P * 25 * 4)) + 1)
    sets = [(d(loc) + 0.001 * i, d(conf) + 0.001 * i) for i in range(nset)]
    sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
    gl = torch.empty_like(sets[0][0]); gcf = torch.empty_like(sets[0][1])
    st = torch.cuda.current_stream().cuda_stream
    sumG = int(off[-1])
    def step(i):
        l, c = sets[i % nset]
        ctx.loss_dev(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, sumG,
                     sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
    for i in range(20): step(i)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): step(i)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / steps * 1e3)
    step(0); torch.cuda.synchronize()
    chk = [int(gl.view(torch.int32).sum(dtype=torch.int64)), int(gcf.view(torch.int32).sum(dtype=torch.int64)),
           sums.cpu().numpy().view(np.int64).tolist()]
    print(json.dumps(dict(B=B, step_us_min=round(min(ts), 2), step_us_med=round(float(np.median(ts)), 2), checksum=chk)))
    ctx.close()


if __name__ == "__main__":
    for B in ([int(a) for a in sys.argv[1:]] or [256, 32]):
        run(B)
