"""Developer timing of the host-buffer call (ssdhead_ctx_multibox_loss_host) beside the raw PCIe copy of its inputs."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR
from objectdetection_ssd_b200.ctx import SSDHeadContext, pinned_empty

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pri = PR.make_priors(); P = pri.shape[0]
ctx = SSDHeadContext(pri.numpy(), max_batch=B)
gb, gc = synth.make_gt(1, B); gx, gcl, off = synth.pack_gt(gb, gc)
loc, conf = synth.make_head(1, B, P)
hl, hc = pinned_empty(loc.shape), pinned_empty(conf.shape); hl[:] = loc; hc[:] = conf
hgl, hgc = pinned_empty(loc.shape), pinned_empty(conf.shape)

def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

dconf = torch.empty(conf.shape, device="cuda")
tconf = torch.from_numpy(hc)
raw = timeit(lambda: dconf.copy_(tconf, non_blocking=True))
def raw_with_memset():
    dconf.copy_(tconf, non_blocking=True); hgc.fill(0); hgl.fill(0)
rawz = timeit(raw_with_memset)
full = timeit(lambda: ctx.loss_host(hl, hc, gx, gcl, off, hgl, hgc))
fwd = timeit(lambda: ctx.loss_host(hl, hc, gx, gcl, off, None, None))
print(json.dumps(dict(B=B, conf_MB=round(conf.nbytes / 1e6, 1), raw_h2d_ms=round(raw, 3), raw_GBs=round(conf.nbytes / raw / 1e6, 1),
                      raw_h2d_plus_single_thread_memset_ms=round(rawz, 3), loss_host_ms=round(full, 3), loss_host_fwd_only_ms=round(fwd, 3))))
ctx.close()
