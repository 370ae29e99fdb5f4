"""Developer tool: per-phase device times of a training-head step (begin = CE stream || match, end = mine)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.ctx import SSDHeadContext
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pri = PR.make_priors(); P = pri.shape[0]
ctx = SSDHeadContext(pri.numpy(), max_batch=B)
gb, gc = synth.make_gt(1, B); gx, gcl, off = synth.pack_gt(gb, gc)
loc, conf = synth.make_head(1, B, P)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tgx, tgc, toff = d(gx), d(gcl), d(off)
sets = [(d(loc) + i, d(conf) + 0.01 * i) for i in range(3)]
sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
gl = torch.empty_like(sets[0][0]); gcf = torch.empty_like(sets[0][1])
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
acc = [0.0, 0.0]; n = 30
for i in range(n + 5):
    l, c = sets[i % 3]
    ev[0].record()
    ctx.loss_begin(c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]), gl.data_ptr(), gcf.data_ptr(), st)
    ev[1].record()
    ctx.loss_end(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, None, sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
    ev[2].record()
    torch.cuda.synchronize()
    if i >= 5:
        acc[0] += ev[0].elapsed_time(ev[1]); acc[1] += ev[1].elapsed_time(ev[2])
# kernels alone, serial, warm
head = MultiboxHead(pri, "cuda")
gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
lib = _lib.load()
ws = head._workspace(_lib.WS_LOSS, B, 0)
tm = tc = 0.0
for i in range(n + 5):
    l, c = sets[i % 3]
    ev[0].record(); m = head.match(gt); ev[1].record()
    lib.ssdhead_ce_stream(c.data_ptr(), B, P, 21, None, gl.data_ptr(), gcf.data_ptr(), ws.data_ptr(), ws.numel(), st)
    ev[2].record(); torch.cuda.synchronize()
    if i >= 5:
        tm += ev[0].elapsed_time(ev[1]); tc += ev[1].elapsed_time(ev[2])
print(json.dumps(dict(B=B, begin_us=round(acc[0] / n * 1e3, 1), end_mine_us=round(acc[1] / n * 1e3, 1),
                      match_alone_us=round(tm / n * 1e3, 1), ce_alone_us=round(tc / n * 1e3, 1))))
