"""Model check of the pipelined sharded exchange's flow control (capi.cu bh_shards_post / collect).

Per rank two in-order streams: S carries the traversal launches (beam_e), X the flag + merge kernels
(signal_e, merge_e). Dependencies, exactly as enqueued by the host code:
    beam_e    after beam_{e-1}; every R/2-th call also after this rank's own merge_{e-R/2}
    signal_e  after beam_e (event E) and after merge_{e-1} (X is in order)
    merge_e   after signal_e and after EVERY rank's signal_e (it spins on the flags)
beam_e of rank A stores into slot e mod R of every rank B. The property that makes the ring safe:
when beam_e starts on any rank, every rank has finished merge_{e-R} (the last reader of that slot).
The test draws random kernel durations (including pathological skew: one rank 50x slower, merges much
slower than traversals) and checks the property for the ring depth the engine uses — and that a ring
WITHOUT the periodic wait, or with a ring that is too shallow for the wait period, does violate it."""
import random

import pytest

RING = 16          # kShardRing in capi.cu


def simulate(nranks, ncalls, ring, wait_every, wait_back, rng, skew):
    """Returns (beam_start[r][e], merge_end[r][e]) under the dependency rules above (1-based calls)."""
    dur_beam = [[rng.uniform(0.5, 1.5) * skew[r] for _ in range(ncalls + 1)] for r in range(nranks)]
    dur_sig = [[rng.uniform(0.001, 0.01) for _ in range(ncalls + 1)] for r in range(nranks)]
    dur_merge = [[rng.uniform(0.01, 3.0) for _ in range(ncalls + 1)] for r in range(nranks)]
    beam_start = [[0.0] * (ncalls + 1) for _ in range(nranks)]
    beam_end = [[0.0] * (ncalls + 1) for _ in range(nranks)]
    sig_end = [[0.0] * (ncalls + 1) for _ in range(nranks)]
    merge_end = [[0.0] * (ncalls + 1) for _ in range(nranks)]
    for e in range(1, ncalls + 1):
        for r in range(nranks):
            t = beam_end[r][e - 1]
            if wait_every and e % wait_every == 0 and e >= ring:
                t = max(t, merge_end[r][e - wait_back])
            beam_start[r][e] = t
            beam_end[r][e] = t + dur_beam[r][e]
        for r in range(nranks):
            sig_end[r][e] = max(beam_end[r][e], merge_end[r][e - 1]) + dur_sig[r][e]
        for r in range(nranks):
            merge_end[r][e] = max(max(sig_end[q][e] for q in range(nranks)), sig_end[r][e]) + dur_merge[r][e]
    return beam_start, merge_end


def violations(nranks, ncalls, ring, beam_start, merge_end):
    bad = 0
    for e in range(ring + 1, ncalls + 1):
        for a in range(nranks):
            for b in range(nranks):
                if beam_start[a][e] < merge_end[b][e - ring]:
                    bad += 1
    return bad


@pytest.mark.parametrize("seed", range(8))
def test_ring_slot_is_never_rewritten_before_its_last_reader_finished(seed):
    rng = random.Random(seed)
    nranks = rng.choice([2, 3, 8])
    skew = [1.0] * nranks
    if seed % 2:
        skew[rng.randrange(nranks)] = 50.0           # one rank far slower than the rest
    if seed % 3 == 0:
        skew = [0.02 * s for s in skew]              # traversals much shorter than merges
    bs, me = simulate(nranks, 200, RING, RING // 2, RING // 2, rng, skew)
    assert violations(nranks, 200, RING, bs, me) == 0


def test_the_periodic_wait_is_what_makes_it_safe():
    rng = random.Random(1)
    skew = [0.02, 0.02, 1.0]                         # fast ranks race ahead of a slow one
    bs, me = simulate(3, 200, RING, 0, 0, rng, skew)            # no wait at all
    assert violations(3, 200, RING, bs, me) > 0
    bs, me = simulate(3, 200, 4, 8, 8, random.Random(1), skew)  # ring shallower than the wait period
    assert violations(3, 200, 4, bs, me) > 0
