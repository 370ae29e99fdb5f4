"""Developer tool: cProfile of the drop-in surface, Losses.ssd((loc, conf), classes, bboxes) + (l1 + l2).backward(),
batch 256 - where the Python / torch time per call goes."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from objectdetection_ssd_b200 import Losses, synth, priors as PR

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
host = len(sys.argv) > 2 and sys.argv[2] == "host"
dev = torch.device("cuda:0")
P = PR.make_priors().shape[0]
gb, gc = synth.make_gt(1, B)
loc, conf = synth.make_head(1, B, P)
to = (lambda a: torch.from_numpy(a)) if host else (lambda a: torch.from_numpy(a).to(dev))
classes, bboxes = [to(c) for c in gc], [to(b) for b in gb]
sets = [((torch.from_numpy(loc).to(dev) + 0.001 * i).requires_grad_(True), (torch.from_numpy(conf).to(dev) + 0.001 * i).requires_grad_(True)) for i in range(3)]

def it(i):
    l, c = sets[i % 3]
    l.grad = None; c.grad = None
    l1, l2 = Losses.ssd((l, c), classes, bboxes)
    (l1 + l2).backward()

for i in range(10): it(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(200): it(i)
torch.cuda.synchronize()
print("wall per call: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
pr = cProfile.Profile(); pr.enable()
for i in range(200): it(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
