"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: image sharding, the packed
float64 all-reduce, global-batch normalisation.  CPU only; per-rank partials come from the
oracle (on the GPU box the same code path carries the kernels' outputs over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maxsquareloss_b200 import dist as mdist
from maxsquareloss_b200 import synth
from oracle import eval_port, loss_math


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partials(lo_all, HW, C, ratio, n_global, gts, prs):
    """What one rank's kernels emit for its images: loss partial normalised by the GLOBAL
    batch, class histogram summed over its images, confusion counts."""
    r = loss_math.fused_iw(lo_all.numpy(), HW, C, ratio)
    n_local = lo_all.shape[0]
    loss_partial = r["loss"] * n_local / n_global           # oracle normalises by n_local
    hist = r["hist"].sum(axis=0)
    cm = eval_port.confusion(gts.numpy(), prs.numpy(), C)
    return loss_partial, hist, cm


def _worker(rank, world, port, n_global, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C, hw, HW = synth.SHAPES["tiny13"]
        lo = synth.head_logits(n_global, C, hw, 11, 3.0)
        gt = synth.blocky_labels(n_global, HW, C, 12, grid=(4, 8))
        pr = synth.noisy_prediction(gt, C, 12)
        a, b = mdist.image_shard(n_global, rank, world)
        buf = torch.zeros(mdist.stats_len(C), dtype=torch.float64)
        if b > a:
            lp, hist, cm = _partials(lo[a:b], HW, C, 0.2, n_global, gt[a:b], pr[a:b])
            mdist.pack_stats(torch.tensor(lp, dtype=torch.float64), torch.from_numpy(hist),
                             torch.from_numpy(cm), out=buf)
        work = mdist.allreduce_stats(buf, async_op=True)
        work.wait()
        loss, hist, cm = mdist.unpack_stats(buf, C)
        q.put((rank, float(loss), hist.numpy(), cm.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global", [(2, 4), (2, 5), (3, 4)])
def test_sharded_equals_single(world, n_global):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, hw, HW = synth.SHAPES["tiny13"]
    lo = synth.head_logits(n_global, C, hw, 11, 3.0)
    gt = synth.blocky_labels(n_global, HW, C, 12, grid=(4, 8))
    pr = synth.noisy_prediction(gt, C, 12)
    ref_loss, ref_hist, ref_cm = _partials(lo, HW, C, 0.2, n_global, gt, pr)
    for _, loss, hist, cm in res:
        assert np.array_equal(hist, ref_hist)               # ints bit-exact
        assert np.array_equal(cm, ref_cm)
        assert abs(loss - ref_loss) <= 1e-12 * abs(ref_loss)


def test_image_shard_covers_batch():
    for n in (1, 2, 7, 8, 500):
        for world in (1, 2, 3, 4, 8):
            blocks = [mdist.image_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip_exact():
    C = 19
    hist = torch.randint(0, 2 ** 40, (C,), dtype=torch.int64)
    cm = torch.randint(0, 2 ** 50, (C, C), dtype=torch.int64)
    buf = mdist.pack_stats(torch.tensor(-0.0123, dtype=torch.float64), hist, cm)
    assert buf.numel() == 8 * 0 + 1 + C + C * C
    loss, h2, cm2 = mdist.unpack_stats(buf, C)
    assert torch.equal(h2, hist) and torch.equal(cm2, cm) and loss.item() == -0.0123
    assert mdist.allreduce_stats(buf) is None                # no process group: no-op
