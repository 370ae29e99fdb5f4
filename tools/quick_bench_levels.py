"""Developer timing: training-head step on per-level tensors vs the concatenated layout (+ what the reference's
Model.py:212-235 pays to build the concatenated tensors: 12 permute+contiguous copies and 2 cats)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.head import MultiboxHead, PackedGT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = 200
pri = PR.make_priors(); P = pri.shape[0]
head = MultiboxHead(pri, "cuda"); lib = _lib.load()
gb, gc = synth.make_gt(1, B)
gt = PackedGT([torch.from_numpy(b) for b in gb], [torch.from_numpy(c) for c in gc], head.dev)
loc, conf = synth.make_head(1, B, P)
grids = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))
nset = 3
sets = []
for i in range(nset):
    l = torch.from_numpy(loc).cuda() + 0.001 * i; c = torch.from_numpy(conf).cuda() + 0.001 * i
    ls, cs, s = [], [], 0
    for hw, a in grids:
        n = hw * hw * a
        ls.append(l[:, s:s + n].contiguous()); cs.append(c[:, s:s + n].contiguous()); s += n
    sets.append((l, c, ls, cs))
gls = [torch.empty_like(t) for t in sets[0][2]]; gcs = [torch.empty_like(t) for t in sets[0][3]]
gl = torch.empty_like(sets[0][0]); gcf = torch.empty_like(sets[0][1])
sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
m = head._match_outputs(gt, False)
ws = head._workspace(_lib.WS_LOSS, B, 0); wm = head._workspace(_lib.WS_MATCH, B, gt.sumG)
st = torch.cuda.current_stream().cuda_stream
structs = []
for (_, _, ls, cs) in sets:
    s_ = _lib.Levels(); s_.num_levels = 6
    for i in range(6):
        s_.count[i] = ls[i].shape[1]; s_.conf[i] = cs[i].data_ptr(); s_.loc[i] = ls[i].data_ptr()
        s_.grad_conf[i] = gcs[i].data_ptr(); s_.grad_loc[i] = gls[i].data_ptr()
    structs.append(s_)

def step_levels(i):
    return lib.ssdhead_multibox_step_levels(ctypes.addressof(structs[i % nset]), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
        head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, gt.sumG, 3, 0.5, sums.data_ptr(), losses.data_ptr(),
        m["cls_u8"].data_ptr(), m["best_prior"].data_ptr(), m["npos"].data_ptr(), ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), st)

def step_cat(i):
    l, c = sets[i % nset][:2]
    return lib.ssdhead_multibox_step(l.data_ptr(), c.data_ptr(), gt.boxes.data_ptr(), gt.classes.data_ptr(), gt.off.data_ptr(),
        head.pri_xyxy.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, gt.sumG, 3, 0.5, sums.data_ptr(), losses.data_ptr(),
        gl.data_ptr(), gcf.data_ptr(), m["cls_u8"].data_ptr(), m["best_prior"].data_ptr(), m["npos"].data_ptr(), None, None,
        ws.data_ptr(), ws.numel(), wm.data_ptr(), wm.numel(), st)

def reference_concat(i):
    # Model.py:212-235 on NCHW conv outputs: permute(0,2,3,1).contiguous().view per map, then two cats
    ls, cs = sets[i % nset][2:]
    return torch.cat([t.view(B, -1, 4) for t in ls], 1), torch.cat([t.view(B, -1, 21) for t in cs], 1)

def timeit(f):
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

_lib.check(step_levels(0), "levels"); _lib.check(step_cat(0), "cat")
print(json.dumps(dict(B=B, step_levels_us=round(timeit(step_levels), 1), step_concatenated_us=round(timeit(step_cat), 1),
                      torch_cat_of_the_levels_us=round(timeit(reference_concat), 1), loss=losses.tolist())))
