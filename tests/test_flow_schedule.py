"""Host-side model of the dataflow schedule of csrc/fused_conv.cuh (flow_conv_fwd_kernel): the item decode and the
dependency rule are restated here and checked by simulation -- every dependency points to a smaller ticket (so the
smallest unfinished ticket can always run: no deadlock with co-resident CTAs), every (phase, row, tile) is issued
exactly once, and a scratch ring slot is never written by pass 1 of a row before pass 3 of its previous tenant is done."""
import heapq
import random

import pytest


def decode(item, tiles, rowctas, dist):
    """ticket -> (phase, row, tile) exactly as the kernel does (rows outside [0, R) are skipped by the caller)."""
    T = 2 * tiles + rowctas
    stage, j = divmod(item, T)
    if j < tiles:
        return 3, stage - 2 * dist, j
    if j < tiles + rowctas:
        return 2, stage - dist, j - tiles
    return 1, stage, j - tiles - rowctas


def deps(phase, row, tiles, rowctas, ring):
    """(counter phase, row, needed count) the item waits for, or None."""
    if phase == 1:
        return (3, row - ring, tiles) if row >= ring else None
    if phase == 2:
        return (1, row, tiles)
    return (2, row, rowctas)


@pytest.mark.parametrize("R,tiles,rowctas,dist,workers", [(16, 4, 4, 2, 7), (5, 8, 4, 1, 3), (40, 16, 16, 2, 37),
                                                           (3, 4, 2, 3, 64), (64, 128, 128, 2, 296)])
def test_schedule_is_complete_ordered_and_ring_safe(R, tiles, rowctas, dist, workers):
    ring = 2 * dist + 2
    T = 2 * tiles + rowctas
    total = (R + 2 * dist) * T
    first_ticket, last_ticket = {}, {}
    seen = set()
    for it in range(total):
        ph, row, tile = decode(it, tiles, rowctas, dist)
        if not 0 <= row < R:
            continue
        assert (ph, row, tile) not in seen
        seen.add((ph, row, tile))
        first_ticket.setdefault((ph, row), it)
        last_ticket[(ph, row)] = it
    assert len(seen) == R * T                                            # every tile of every phase of every row, once
    for (ph, row), t0 in first_ticket.items():                           # dependencies point backwards in ticket order
        d = deps(ph, row, tiles, rowctas, ring)
        if d:
            assert last_ticket[(d[0], d[1])] < t0

    # event simulation with random durations: workers take tickets in order; an item starts once the counter it
    # waits for is full (all producer tiles finished)
    rng = random.Random(1)
    count, done_at, full_at = {}, {}, {}
    free = [(0.0, w) for w in range(workers)]
    heapq.heapify(free)
    tenant = {}                                                          # ring slot -> row that owns it
    for it in range(total):
        ph, row, tile = decode(it, tiles, rowctas, dist)
        t_free, w = heapq.heappop(free)
        if not 0 <= row < R:
            heapq.heappush(free, (t_free, w))
            continue
        start = t_free
        d = deps(ph, row, tiles, rowctas, ring)
        if d:
            assert (d[0], d[1]) in full_at, "a producer has a larger ticket: the wait could deadlock"
            start = max(start, full_at[(d[0], d[1])])
        if ph == 1:                                                      # ring safety: the previous tenant is drained
            slot = row % ring
            prev = tenant.get(slot)
            if prev is not None and prev != row:
                assert full_at[(3, prev)] <= start
            tenant[slot] = row
        end = start + rng.uniform(0.5, 2.0)
        key = (ph, row)
        count[key] = count.get(key, 0) + 1
        done_at[key] = max(done_at.get(key, 0.0), end)
        if count[key] == (rowctas if ph == 2 else tiles):
            full_at[key] = done_at[key]
        heapq.heappush(free, (end, w))
    assert len(full_at) == 3 * R
