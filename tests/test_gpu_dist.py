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


def _worker_box(rank, world, port, q):
    """One-call sharded steps with the statistics exchanged through the NVLink peer-memory mailboxes (the exchange rides in
    the steps' finalisation kernels) and, for comparison, through ncclAllReduce on a second communicator."""
    sys.path.insert(0, ROOT)
    import datetime
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    from maxsquareloss_b200 import _lib, dist as mdist, synth
    lib = _lib.load()
    C, (h, w), (H, W), N, K = 19, (33, 65), (257, 513), 2, 13
    lay = _lib.state_layout(N, C)
    los = [synth.head_logits(N * world, C, (h, w), 100 + i, 4.0)[rank * N:(rank + 1) * N].cuda() for i in range(K)]
    stream = torch.cuda.current_stream().cuda_stream

    def run(comm, flush_at):
        accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
        outs = [torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda") for _ in range(K)]     # one per step in flight
        aux = torch.empty(lib.msq_fused_aux_bytes(N, H, W), dtype=torch.uint8, device="cuda")
        grad = torch.empty_like(los[0])
        red = torch.zeros(K, 1 + C, dtype=torch.float64, device="cuda")
        one = outs[0]                      # ONE `out` reused by every step: the communicator keeps no pointer into it
        fetched = -1
        for i in range(K):
            _lib.check(lib.msq_fused_fwd_bwd(_lib.MODE_IW, los[i].data_ptr(), N, C, h, w, H, W, 0.2, N * world, accum.data_ptr(),
                                             one.data_ptr(), aux.data_ptr(), None, 0.1, grad.data_ptr(), comm._h, 0, stream))
            if i in flush_at:
                comm.join()
                for j in range(fetched + 1, i + 1):
                    comm.result(1 + C, lag=i - j, out=red[j])
                fetched = i
            elif not comm.peer_memory:                        # NCCL: the fetch waits (on the stream) for the step's collective
                comm.result(1 + C, lag=0, out=red[i])
                fetched = i
            elif i - 2 > fetched:                             # mailboxes: vector i-2 exists once step i has been enqueued
                comm.result(1 + C, lag=2, out=red[i - 2])
                fetched = i - 2
        comm.join()
        for j in range(fetched + 1, K):
            comm.result(1 + C, lag=K - 1 - j, out=red[j])
        torch.cuda.synchronize()
        local = one[lay.stats_off:lay.stats_off + 8 * (1 + C)].view(torch.float64).cpu().numpy()
        assert local[1:].sum() == N * H * W                  # the caller's `out` keeps the rank-LOCAL statistics
        return red.cpu().numpy()

    box, nccl = mdist.StatsComm(), mdist.StatsComm(peer_memory=False)
    got = [run(box, ()), run(box, (0, 1, 6)), run(box, (K - 2,))]        # the communicator is reused; flushes anywhere
    ref = run(nccl, ())
    q.put((rank, box.peer_memory, nccl.peer_memory, box.errors(), got, ref))
    box.close()
    nccl.close()
    dist.destroy_process_group()


def test_peer_memory_mailboxes_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_box, args=(r, 2, 29534, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert not res[0][2] and not res[1][2]
    if not (res[0][1] and res[1][1]):
        pytest.skip("CUDA IPC peer mapping is not available on this box: the communicator stayed on NCCL")
    for rank, _, _, err, got, ref in res:
        assert err == 0
        assert ref[:, 1:].sum() == 13 * 4 * 257 * 513              # every pixel of both ranks' images counted, every step
        for g in got:
            assert np.array_equal(g, ref)                          # two ranks: a + b in rank order == NCCL's sum, bit for bit
    assert np.array_equal(res[0][4][0], res[1][4][0])              # and identical on both ranks


def _worker_ce(rank, world, port, q):
    """The sharded cross-entropy mean on REAL ranks: MultiLevelTargetLoss with the library communicator (exact uint64
    all-reduce of [ce_fix | nvalid]) and CrossEntropyLoss2d with a torch.distributed group (int64 all_reduce of the same pair)."""
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
    C, hw, HW, NG = 13, (9, 17), (64, 128), 4
    a, b = mdist.image_shard(NG, rank, world)
    lo1 = synth.head_logits(NG, C, hw, 61, 4.0)
    lo2 = synth.second_head(lo1, 61)
    y = synth.blocky_labels(NG, HW, C, 62, grid=(4, 8))
    x1, x2 = lo1[a:b].cuda().requires_grad_(True), lo2[a:b].cuda().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, C, 0.2)
    crit.global_batch = NG
    multi = msq.MultiLevelTargetLoss(crit, threshold=0.9, lambda_target=0.1, lambda_seg=0.1, group=comm)
    l1, l2 = multi((x1, x2), HW)
    (l1 + l2).backward()
    xs = lo1[a:b].cuda().requires_grad_(True)
    ce = msq.CrossEntropyLoss2d()                       # group=None: the default torch.distributed group
    ls = ce(xs, y[a:b].cuda())
    ls.backward()
    torch.cuda.synchronize()
    q.put((rank, (a, b), l2.item(), int(multi.last_nvalid.item()), x2.grad.cpu().numpy(), ls.item(), int(ce.last_nvalid.item()),
           xs.grad.cpu().numpy()))
    comm.close()
    dist.destroy_process_group()


def test_sharded_cross_entropy_means_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import maxsquareloss_b200 as msq
    from maxsquareloss_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_ce, args=(r, 2, 29535, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    C, hw, HW, NG = 13, (9, 17), (64, 128), 4
    lo1 = synth.head_logits(NG, C, hw, 61, 4.0)
    lo2 = synth.second_head(lo1, 61)
    y = synth.blocky_labels(NG, HW, C, 62, grid=(4, 8))
    x1, x2 = lo1.cuda().requires_grad_(True), lo2.cuda().requires_grad_(True)
    multi = msq.MultiLevelTargetLoss(msq.IW_MaxSquareloss(-1, C, 0.2), threshold=0.9, lambda_target=0.1, lambda_seg=0.1, group=False)
    l1, l2 = multi((x1, x2), HW)
    (l1 + l2).backward()
    xs = lo1.cuda().requires_grad_(True)
    ce = msq.CrossEntropyLoss2d(group=False)
    ls = ce(xs, y.cuda())
    ls.backward()
    for rank, (a, b), l2r, nv, g2, lsr, nvs, gs in res:
        assert nv == int(multi.last_nvalid.item()) and nvs == int(ce.last_nvalid.item())          # global counts: exact
        assert abs(l2r - l2.item()) <= 1e-6 * abs(l2.item()) and abs(lsr - ls.item()) <= 1e-6 * abs(ls.item())
        r2, rs = x2.grad[a:b].cpu().numpy(), xs.grad[a:b].cpu().numpy()
        assert np.abs(g2 - r2).max() <= 1e-5 * np.abs(r2).max()
        assert np.abs(gs - rs).max() <= 1e-5 * np.abs(rs).max()


def _worker_lost(rank, world, port, q):
    """A peer that stops stepping: the surviving rank's mailbox reductions time out (0.5 s here), its statistics become
    NaN, the error is reported -- and nothing hangs."""
    sys.path.insert(0, ROOT)
    import datetime
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    from maxsquareloss_b200 import _lib, dist as mdist, synth
    lib = _lib.load()
    C, (h, w), (H, W), N = 19, (33, 65), (257, 513), 2
    lay = _lib.state_layout(N, C)
    lo = synth.head_logits(N, C, (h, w), 7 + rank, 4.0).cuda()
    comm = mdist.StatsComm()
    if not comm.peer_memory:
        q.put((rank, "no-ipc"))
        dist.destroy_process_group()
        return
    comm.set_timeout(0.5)
    accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
    out = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda")
    aux = torch.empty(lib.msq_fused_aux_bytes(N, H, W), dtype=torch.uint8, device="cuda")
    grad = torch.empty_like(lo)
    stream = torch.cuda.current_stream().cuda_stream
    steps = 6 if rank == 0 else 3                        # rank 1 "dies" after three steps
    codes = []
    t0 = time.perf_counter()
    for i in range(steps):
        codes.append(lib.msq_fused_fwd_bwd(_lib.MODE_IW, lo.data_ptr(), N, C, h, w, H, W, 0.2, N * world, accum.data_ptr(),
                                           out.data_ptr(), aux.data_ptr(), None, 0.1, grad.data_ptr(), comm._h, 0, stream))
        torch.cuda.synchronize()
    rc_join = lib.msq_comm_join(comm._h, 0, stream) if rank == 0 else 0        # the dead rank does not even flush
    torch.cuda.synchronize()
    took = time.perf_counter() - t0
    red = comm.result(1 + C, lag=0).cpu().numpy() if rank == 0 else None
    if rank == 0:
        codes.append(lib.msq_comm_join(comm._h, 0, stream))
    q.put((rank, codes, rc_join, comm.errors(), took, red))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_lost_peer_times_out_and_is_reported():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from maxsquareloss_b200 import _lib
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_lost, args=(r, 2, 29536, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    if res[0][1] == "no-ipc":
        pytest.skip("CUDA IPC peer mapping is not available on this box")
    _, codes, rc_join, err, took, red = res[0]
    assert took < 30.0                                              # bounded: a few 0.5 s time-outs, never a hang
    assert err != 0                                                 # the loss is recorded ...
    assert _lib.load().msq_error_string(-6).startswith(b"msq:")
    assert -6 in codes + [rc_join]                                  # ... and returned (MSQ_E_PEER) by a later call
    assert np.isnan(red).any()                                      # the vector that needed the dead rank is NaN, not garbage
    assert res[1][3] == 0                                           # the rank that stopped saw no loss itself
