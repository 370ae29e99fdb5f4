"""Developer timing of the stress configuration (BASELINE.json configs[4]): SSD512-style 24 564 priors, 100 gt boxes per
image, batch 128 per GPU: match + loss step, and detect (decode + NMS)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pri = PR.make_priors(PR.SSD512_SPEC); P = pri.shape[0]
head = MultiboxHead(pri, "cuda"); lib = _lib.load()
gb, gc = synth.make_gt(4, B, G, G)
gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
loc, conf = synth.make_head(4, B, P)
l, c = torch.from_numpy(loc).cuda(), torch.from_numpy(conf).cuda()
gl, gcf = torch.empty_like(l), torch.empty_like(c)
sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
m = head._match_outputs(gt, False)
ws = head._workspace(_lib.WS_LOSS, B, 0); wm = head._workspace(_lib.WS_MATCH, B, gt.sumG)
st = torch.cuda.current_stream().cuda_stream

def step(i):
    return lib.ssdhead_multibox_step(l.data_ptr(), c.data_ptr(), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
        head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, gt.sumG, 3, 0.5, sums.data_ptr(), losses.data_ptr(),
        gl.data_ptr(), gcf.data_ptr(), m["cls_u8"].data_ptr(), m["best_prior"].data_ptr(), m["npos"].data_ptr(), None, None,
        ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), st)

dl, dc = synth.make_head(5, B, P, loc_scale=0.5, bg_bias=6.0)
dl, dc = torch.from_numpy(dl).cuda(), torch.from_numpy(dc).cuda()
top_k = 200
ob = torch.empty(B, top_k, 4, device="cuda"); op = torch.empty(B, top_k, device="cuda")
oc = torch.empty(B, top_k, dtype=torch.int32, device="cuda"); oi = torch.empty_like(oc); on = torch.empty(B, dtype=torch.int32, device="cuda")
wd = torch.zeros(int(lib.ssdhead_workspace_bytes(_lib.WS_DETECT, B, P, 21, 0)) + 256, dtype=torch.uint8, device="cuda")

def det(i):
    return lib.ssdhead_detect(dl.data_ptr(), dc.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, 0.01, 0.45, top_k, None, 0,
                              ob.data_ptr(), op.data_ptr(), oc.data_ptr(), oi.data_ptr(), on.data_ptr(), wd.data_ptr(), wd.numel(), st)

def timeit(f, steps=100):
    for i in range(3): _lib.check(f(i), "call")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

t1, t2 = timeit(step), timeit(det)
print(json.dumps(dict(B=B, P=P, gt_per_image=G, loss_step_us=round(t1, 1), loss_img_per_s=round(B / (t1 * 1e-6)),
                      loss_roofline=round(B * P * 25 * 4 * 2 / (t1 * 1e-6) / 1e9 / 6538.6, 3),
                      detect_step_us=round(t2, 1), detect_img_per_s=round(B / (t2 * 1e-6)), npos=int(m["npos"][-1]),
                      losses=losses.tolist(), det_cnt=on[:3].tolist())))
