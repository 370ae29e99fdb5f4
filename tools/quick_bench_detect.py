"""Developer timing loop for the detect path (device-resident inputs, CUDA events, direct C-ABI calls)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.head import MultiboxHead, detect

def run(B, bias, steps=50, warm=5, top_k=200):
    pri = PR.make_priors(); P = pri.shape[0]
    head = MultiboxHead(pri, "cuda")
    lib = _lib.load()
    loc, conf = synth.make_head(3, B, P, loc_scale=0.5, bg_bias=bias)
    nset = max(2, int(300e6 // (B * P * 25 * 4)) + 1)
    sets = [(torch.from_numpy(loc).cuda() + 0.001 * i, torch.from_numpy(conf).cuda()) for i in range(nset)]
    ob = torch.empty(B, top_k, 4, device="cuda"); op = torch.empty(B, top_k, device="cuda")
    oc = torch.empty(B, top_k, dtype=torch.int32, device="cuda"); oi = torch.empty_like(oc)
    on = torch.empty(B, dtype=torch.int32, device="cuda")
    ws = torch.zeros(int(lib.ssdhead_workspace_bytes(_lib.WS_DETECT, B, P, 21, 0)) + 256, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def step(i):
        l, c = sets[i % nset]
        return lib.ssdhead_detect(l.data_ptr(), c.data_ptr(), head.pri_cxcywh.data_ptr(), B, P, 21, 0.01, 0.45, top_k, None, 0,
                                  ob.data_ptr(), op.data_ptr(), oc.data_ptr(), oi.data_ptr(), on.data_ptr(),
                                  ws.data_ptr(), ws.numel(), st)
    for i in range(warm): _lib.check(step(i), "ssdhead_detect")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): step(i)
    e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1) / steps
    import ctypes
    nfb = ctypes.c_int32(-1)
    lib.ssdhead_detect_fallbacks(ws.data_ptr(), ws.numel(), B, P, 21, 0, ctypes.addressof(nfb), st)
    print(json.dumps(dict(B=B, bias=bias, fallbacks=nfb.value, step_us=round(tt * 1e3, 1), img_per_s=round(B / (tt * 1e-3)),
                          frac_of_6538=round(B * 878800 / (tt * 1e-3) / 1e9 / 6538.6, 4), cnt=on[:4].tolist())))

if __name__ == "__main__":
    for bias in (8.0, 6.0, 4.0):
        run(64, bias)
    for a in sys.argv[1:]:
        run(int(a), 6.0)
