"""Model check of the peer-memory mailbox protocol (maxsquareloss_b200/csrc/common.cuh, PeerBox; comm.cu): CPU only.

Every rank runs, in stream order, one exchange per step i (a second CTA of that step's finalisation kernel):
    push   vector i-1 into slot (i-1) % S of EVERY rank's ring          (fire and forget)
    reduce vector i-2: spin until all ranks' cells of slot (i-2) % S carry sequence number i-2, then sum them
and msq_comm_join: push the newest vector, reduce the one before it, reduce the newest.  The model interleaves the
ranks at random (any rank whose next operation is not blocked may run) and checks that no rank deadlocks, that a cell
is never overwritten before every reader that still needs it has read it, and that every reduced vector is the sum of
what the ranks produced for that step.  It also shows that a ring of 3 slots is NOT enough (the library uses 8)."""
import os
import random
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def simulate(world, steps, slots, seed, flush_at=()):
    rng = random.Random(seed)
    # ring[dst][slot][src] = (seq, value)
    ring = [[[(0, None)] * world for _ in range(slots)] for _ in range(world)]
    produced = lambda r, s: 1000 * s + r                       # rank r's statistic of step s
    reduced = [dict() for _ in range(world)]                    # rank -> {seq: sum}
    consumed = [0] * world                                      # highest seq rank r has finished reducing

    def program(r):                                             # generator of blocking operations of rank r
        last_pushed, last_reduced = 0, 0
        for i in range(1, steps + 1):
            if i - 1 >= 1 and i - 1 > last_pushed:
                yield ("push", i - 1)
                last_pushed = i - 1
            if i - 2 >= 1 and i - 2 > last_reduced:
                yield ("reduce", i - 2)
                last_reduced = i - 2
            if i in flush_at or i == steps:                     # msq_comm_join after step i
                if i > last_pushed:
                    yield ("push", i)
                    last_pushed = i
                for s in range(last_reduced + 1, i + 1):
                    yield ("reduce", s)
                last_reduced = i
    progs = [program(r) for r in range(world)]
    pending = [next(p, None) for p in progs]
    while any(op is not None for op in pending):
        runnable = []
        for r, op in enumerate(pending):
            if op is None:
                continue
            if op[0] == "push" or all(ring[r][op[1] % slots][p][0] == op[1] for p in range(world)):
                runnable.append(r)
        assert runnable, f"deadlock: {pending}"
        r = rng.choice(runnable)
        kind, s = pending[r]
        if kind == "push":
            for dst in range(world):
                old_seq = ring[dst][s % slots][r][0]
                # the cell being overwritten must not be needed any more by its owner
                assert old_seq == 0 or consumed[dst] >= old_seq, \
                    f"rank {r} overwrites seq {old_seq} in rank {dst}'s ring (it has only consumed {consumed[dst]})"
                ring[dst][s % slots][r] = (s, produced(r, s))
        else:
            reduced[r][s] = sum(ring[r][s % slots][p][1] for p in range(world))
            consumed[r] = s
        pending[r] = next(progs[r], None)
    for r in range(world):
        for s in range(1, steps + 1):
            assert reduced[r][s] == sum(produced(p, s) for p in range(world)), (r, s)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_protocol_is_safe_with_the_library_ring(world):
    with open(os.path.join(ROOT, "maxsquareloss_b200", "csrc", "common.cuh")) as f:
        slots = int(re.search(r"constexpr int kBoxSlots = (\d+);", f.read()).group(1))
    assert slots >= 4
    for seed in range(40):
        simulate(world, 40, slots, seed)
        simulate(world, 25, 4, seed, flush_at=(1, 2, 7, 8, 20))       # 4 slots is the minimum that is safe


def test_three_slots_are_not_enough():
    with pytest.raises(AssertionError, match="overwrites"):
        for seed in range(200):
            simulate(4, 30, 3, seed)
