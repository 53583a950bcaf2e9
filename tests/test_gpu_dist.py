"""Two-rank NCCL test of the library's own statistics all-reduce (msq_comm_*, dist.StatsComm) and of the image-sharded
loss: needs >= 2 GPUs (skipped otherwise).  Spawns two processes, one per GPU, rendezvous on 127.0.0.1."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import datetime
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    import maxsquareloss_b200 as msq
    from maxsquareloss_b200 import dist as mdist, synth
    comm = mdist.StatsComm()
    C, hw, HW = 19, (33, 65), (257, 513)
    lo = synth.head_logits(4, C, hw, 5, 4.0)
    a, b = mdist.image_shard(4, rank, world)
    crit = msq.IW_MaxSquareloss(-1, C, 0.2)
    crit.global_batch = 4
    x = lo[a:b].cuda().requires_grad_(True)
    loss = crit(x, out_size=HW)
    stats = crit.last_stats.clone()
    comm.allreduce(stats)                  # side stream, overlaps the backward
    loss.backward()
    comm.join()
    for _ in range(50):                    # many back-to-back collectives: ordering on the side stream
        comm.allreduce(stats)
    comm.join()
    torch.cuda.synchronize()
    q.put((rank, stats.cpu().numpy() / (world ** 50), x.grad.cpu().numpy(), (a, b)))
    comm.close()
    dist.destroy_process_group()


def test_stats_allreduce_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import maxsquareloss_b200 as msq
    from maxsquareloss_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    C, hw, HW = 19, (33, 65), (257, 513)
    lo = synth.head_logits(4, C, hw, 5, 4.0)
    x = lo.cuda().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, C, 0.2)
    whole = crit(x, out_size=HW)
    whole.backward()
    ref = crit.last_stats.cpu().numpy()
    for rank, stats, grad, (a, b) in res:
        assert np.array_equal(stats[1:], ref[1:])                       # class histogram: exact
        assert abs(stats[0] - ref[0]) <= 1e-8 * abs(ref[0])              # loss: per-thread fp32 run sums depend on the row partition
        g = x.grad[a:b].cpu().numpy()
        assert np.abs(grad - g).max() <= 1e-5 * np.abs(g).max()          # no exchange needed for dL/dlogits
