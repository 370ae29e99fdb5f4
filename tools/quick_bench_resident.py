"""Developer timing loop: the training-head step with dense gradient writes against resident gradient tensors."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.ctx import SSDHeadContext

def run(B, steps=int(os.environ.get("STEPS", 300)), warm=int(os.environ.get("WARM", 10))):
    pri = PR.make_priors(); P = pri.shape[0]
    ctx = SSDHeadContext(pri.numpy(), max_batch=B)
    nset = max(2, int(300e6 // (B * P * 25 * 4)) + 1)
    sets = []
    for i in range(nset):
        gb, gc = synth.make_gt(100 + i, B)
        loc, conf = synth.make_head(100 + i, B, P)
        gx, gcl, off = synth.pack_gt(gb, gc)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        sets.append((d(loc), d(conf), d(gx), d(gcl), d(off), int(off[-1])))
    gl = torch.empty(B, P, 4, device="cuda"); gc_ = torch.empty(B, P, 21, device="cuda")
    sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for mode in ("dense", "resident"):
        def step(i, fresh=False):
            l, c, gx, gcl, off, sg = sets[i % nset]
            if mode == "dense":
                ctx.loss_dev(l.data_ptr(), c.data_ptr(), gx.data_ptr(), gcl.data_ptr(), off.data_ptr(), B, sg,
                             sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gc_.data_ptr(), st)
            else:
                ctx.loss_dev_resident(l.data_ptr(), c.data_ptr(), gx.data_ptr(), gcl.data_ptr(), off.data_ptr(), B, sg,
                                      sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gc_.data_ptr(), st, fresh=fresh)
        step(0, fresh=True)
        for i in range(warm): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps): step(i)
        e1.record(); torch.cuda.synchronize()
        out[mode] = round(e0.elapsed_time(e1) / steps * 1e3, 1)
        out[mode + "_chk"] = float(gc_.double().abs().sum())
    print(json.dumps(dict(B=B, **out)))
    ctx.close()

if __name__ == "__main__":
    for B in [int(a) for a in sys.argv[1:]] or [256, 32]:
        run(B)
