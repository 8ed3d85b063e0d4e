"""Model check of the peer-memory halo protocol (csrc/comm.cu, JETCU_P2P_HALO=1) — the DESIGN, not the CUDA code.

Each rank runs, in program order and for every exchange m = 1, 2, ...:
    push_write    store its boundary planes into slot m & 1 of each neighbour's mailbox        (k_p2p_push, all CTAs)
    push_publish  publish message number m in each neighbour's flag                            (k_p2p_push, last CTA)
    pull_wait     block until every neighbour's flag in the own mailbox is >= m                (k_p2p_pull, thread 0)
    pull_read     copy the halo out of slot m & 1 of the own mailbox                           (k_p2p_pull, all CTAs)
There are no credits: the claim is that two slots per direction suffice because a rank can only reach push_write(m) after
its pull of m - 1 completed, i.e. after the neighbour published m - 1, which the neighbour did (program order) after its own
pull_read(m - 2) — the previous user of the slot.  This test explores EVERY interleaving of those steps across the ranks of a
chain (the GPUs run asynchronously; any interleaving that respects program order and the waits can happen) and checks that
every pull_read sees exactly message m of the right neighbour, that the system never deadlocks, and — as a control — that the
same protocol with ONE slot per direction is caught overwriting unread data."""
import pytest


def explore(world, exchanges, slots):
    """DFS over all reachable states.  Returns (states visited, first violation or None)."""
    steps = 4 * exchanges                                   # per-rank program: 4 steps per exchange
    neighbours = [[q for q in (r - 1, r + 1) if 0 <= q < world] for r in range(world)]
    # state: (pc per rank, mailbox[(dst, src, slot)] = message number stored, flag[(dst, src)] = published number)
    boxes = {(d, s, k): 0 for d in range(world) for s in neighbours[d] for k in range(slots)}
    flags = {(d, s): 0 for d in range(world) for s in neighbours[d]}
    keys_b, keys_f = sorted(boxes), sorted(flags)
    start = (tuple([0] * world), tuple(boxes[k] for k in keys_b), tuple(flags[k] for k in keys_f))
    seen = {start}
    stack = [start]
    while stack:
        pcs, bvals, fvals = stack.pop()
        b = dict(zip(keys_b, bvals)); f = dict(zip(keys_f, fvals))
        moved = False
        for r in range(world):
            pc = pcs[r]
            if pc == steps:
                continue
            m, phase = pc // 4 + 1, pc % 4
            nb, nf = dict(b), dict(f)
            if phase == 0:                                  # push_write
                for q in neighbours[r]:
                    nb[(q, r, m % slots)] = m
            elif phase == 1:                                # push_publish
                for q in neighbours[r]:
                    nf[(q, r)] = m
            elif phase == 2:                                # pull_wait
                if any(f[(r, q)] < m for q in neighbours[r]):
                    continue                                # blocked
            else:                                           # pull_read
                for q in neighbours[r]:
                    if b[(r, q, m % slots)] != m:
                        return len(seen), "rank %d read message %d of rank %d while expecting %d" % (r, b[(r, q, m % slots)], q, m)
            moved = True
            npcs = list(pcs); npcs[r] += 1
            nxt = (tuple(npcs), tuple(nb[k] for k in keys_b), tuple(nf[k] for k in keys_f))
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved and any(pc != steps for pc in pcs):
            return len(seen), "deadlock at %s" % (pcs,)
    return len(seen), None


@pytest.mark.parametrize("world", [2, 3, 4])
def test_two_slots_need_no_credits(world):
    states, bad = explore(world, exchanges=5 if world < 4 else 4, slots=2)
    assert bad is None, bad
    assert states > 100


def test_one_slot_is_caught():
    """Control: with a single slot a fast neighbour overwrites message m before it is read — the checker must see it."""
    _, bad = explore(2, exchanges=3, slots=1)
    assert bad is not None and "read message" in bad
