"""Developer tool: SM-clock phase times of the first image's mining CTA (mine_kernel).

Needs a developer build:  SSDHEAD_NVCC_EXTRA=-DSSDHEAD_PHASE_TIMES python -m objectdetection_ssd_b200.build --force
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from objectdetection_ssd_b200 import synth, priors as PR, _lib
from objectdetection_ssd_b200.ctx import SSDHeadContext

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # > 0: the stress configuration (SSD512-style priors, G gts per image)
pri = PR.make_priors(PR.SSD512_SPEC) if G else PR.make_priors(); P = pri.shape[0]
ctx = SSDHeadContext(pri.numpy(), max_batch=B, max_total_gt=B * max(G, 10))
gb, gc = synth.make_gt(1, B, G, G) if G else synth.make_gt(1, B); gx, gcl, off = synth.pack_gt(gb, gc)
loc, conf = synth.make_head(1, B, P)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tgx, tgc, toff = d(gx), d(gcl), d(off)
sets = [(d(loc) + i, d(conf) + 0.01 * i) for i in range(3)]
sums = torch.empty(2, dtype=torch.float64, device="cuda"); losses = torch.empty(2, device="cuda")
gl = torch.empty_like(sets[0][0]); gcf = torch.empty_like(sets[0][1])
st = torch.cuda.current_stream().cuda_stream
lib = _lib.load()
import ctypes as _ct
gm = (_ct.c_ulonglong * 8)()
names = ["keys (natural classes)", "finalise + patch", "select+list", "wait total", "grad rows", "publish"]
for it in range(5):
    l, c = sets[it % 3]
    lib.ssdhead_debug_gmarks(gm, 1)
    ctx.loss_dev(l.data_ptr(), c.data_ptr(), tgx.data_ptr(), tgc.data_ptr(), toff.data_ptr(), B, int(off[-1]),
                 sums.data_ptr(), losses.data_ptr(), gl.data_ptr(), gcf.data_ptr(), st)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.ssdhead_debug_phases_loss(buf)
    ph = list(buf)[:7]
    lib.ssdhead_debug_gmarks(gm, 0)
    g = list(gm)
    print("   global timer: stream end -> first mining CTA past wait %.2f us, -> last CTA past wait %.2f us, mining (first start -> last end) %.2f us" % ((g[1] - g[0]) / 1e3, (g[2] - g[0]) / 1e3, (g[3] - g[1]) / 1e3))
    print(it, " ".join(f"{names[i]}={(ph[i + 1] - ph[i]) / 1965.0:.2f}us" for i in range(6)), f"total={(ph[6] - ph[0]) / 1965.0:.2f}us")
import numpy as _np
buf = (_ct.c_longlong * (12 * B))()
lib.ssdhead_debug_cta(buf, B)
a = _np.array(list(buf), dtype=_np.float64).reshape(B, 12)
us = a[:, :3] / 1965.0
order = _np.argsort(us[:, 2])
print("per-CTA time from its wait to its end (us): min %.1f  median %.1f  p90 %.1f  max %.1f" % (us[:, 2].min(), _np.median(us[:, 2]), _np.percentile(us[:, 2], 90), us[:, 2].max()))
print("  slowest CTAs: " + ", ".join("b=%d nsel=%d sel=%.1f rows=%.1f end=%.1f" % (i, a[i, 3], us[i, 0], us[i, 1], us[i, 2]) for i in order[-5:]))
print("  slowest CTAs, us from the CTA's start: " + "; ".join("b=%d nsel=%d total-known=%.1f trip0[w0 conf %.1f, w0 end %.1f, w15 end %.1f] trip1[w0 conf %.1f, w0 end %.1f, w15 end %.1f] rows=%.1f" % (
    i, a[i, 3], a[i, 4] / 1965, a[i, 9] / 1965, a[i, 5] / 1965, a[i, 7] / 1965, a[i, 10] / 1965, a[i, 6] / 1965, a[i, 8] / 1965, us[i, 1]) for i in order[-6:]))
print("  fastest CTAs: " + ", ".join("b=%d nsel=%d sel=%.1f rows=%.1f end=%.1f" % (i, a[i, 3], us[i, 0], us[i, 1], us[i, 2]) for i in order[:3]))
print("  corr(nsel, end) = %.2f; mean end for b < 108 (2 CTAs/SM partner of b+148): %.1f, 108 <= b < 148 (alone): %.1f, b >= 148: %.1f" % (
    _np.corrcoef(a[:, 3], us[:, 2])[0, 1], us[:108, 2].mean(), us[108:148, 2].mean() if B > 148 else float("nan"), us[148:, 2].mean() if B > 148 else float("nan")))

