"""CPU, world_size 2 over gloo: the sharded-loss arithmetic.  Each rank evaluates its image
shard (here with the oracle standing in for the kernels), all-reduces the positive count and the two loss sums
with objectdetection_ssd_b200.dist, and must reproduce the single-call loss on the full batch."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from objectdetection_ssd_b200 import synth
from oracle import ssd_oracle as O


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    from objectdetection_ssd_b200.dist import shard_lists, allreduce_loss_parts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        pri = O.make_priors()
        gb, gc = synth.make_gt(17, B)
        loc, conf = synth.make_head(17, B, pri.shape[0])
        tb = [torch.from_numpy(b) for b in gb]
        tc = [torch.from_numpy(c) for c in gc]
        sl, sc, sb, sk = shard_lists(rank, world, torch.from_numpy(loc), torch.from_numpy(conf), tb, tc)
        r = O.multibox_loss(sl, sc, sb, sk, pri)
        n = float(r["npos_total"])
        sum_l1 = r["loc_loss"].double() * 4.0 * n           # undo the local normalisation -> partial sums
        sum_ce = r["conf_loss"].double() * n
        npos = torch.tensor([r["npos_total"]], dtype=torch.int64)
        l1, l2 = allreduce_loss_parts(sum_l1, sum_ce, npos)
        q.put((rank, float(l1), float(l2), int(npos)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_loss_equals_full_batch():
    B, world, port = 6, 2, 29541 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pri = O.make_priors()
    gb, gc = synth.make_gt(17, B)
    loc, conf = synth.make_head(17, B, pri.shape[0])
    full = O.multibox_loss(torch.from_numpy(loc), torch.from_numpy(conf), [torch.from_numpy(b) for b in gb],
                           [torch.from_numpy(c) for c in gc], pri)
    for rank, l1, l2, npos in res:
        assert npos == full["npos_total"]
        assert abs(l1 - full["loc_loss"].item()) <= 1e-5 * abs(full["loc_loss"].item())
        assert abs(l2 - full["conf_loss"].item()) <= 1e-5 * abs(full["conf_loss"].item())
