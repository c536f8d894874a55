"""Model check of the tensor-parallel DECODE all-reduce through mailboxes (csrc/llm.cu: llm_reduce_rms_kernel, MAIL form).
Protocol, per exchange point k (one kernel launch per rank and exchange point): a rank stores its partial row into the PEER's
mailbox of parity k & 1, polls its OWN mailbox of parity k & 1 until every slot holds a value instead of the sentinel, uses the
values and empties the slots.  There is no flag: the value is its own arrival flag.
The simulation runs the two ranks as state machines under random schedules with random NVLink delivery delays (stores to
different slots may land in any order) and checks what the kernel relies on: a store always lands in an EMPTY slot, a poll never
consumes a value of another exchange point, and the ranks never drift more than one exchange point apart -- which is why two
parities are enough.  Stores of one kernel are performed before the rank's next kernel starts (kernel boundary), modelled as:
a rank's reset of exchange k is visible before anything it stores in exchange k + 1 is issued."""
import random

import pytest

EMPTY = None
SLOTS = 4        # slots per row stand-in: the kernel has H / 4 sixteen-byte slots per row, each polled by one thread


class Rank(object):
    def __init__(self):
        self.mail = [[EMPTY] * SLOTS, [EMPTY] * SLOTS]
        self.k = 0                 # exchange point this rank is working on
        self.phase = "push"        # push -> poll -> (reset, next k)
        self.consumed = []


def run(seed, n_exchanges=40):
    rng = random.Random(seed)
    ranks = [Rank(), Rank()]
    in_flight = []                 # (deliver_at, dst_rank, parity, slot, (k, value))
    now = 0
    max_drift = 0
    while any(r.k < n_exchanges for r in ranks):
        now += 1
        # deliver stores whose time has come, in random order among those due
        due = [w for w in in_flight if w[0] <= now]
        rng.shuffle(due)
        for w in due:
            in_flight.remove(w)
            _, dst, parity, slot, payload = w
            assert ranks[dst].mail[parity][slot] is EMPTY, "a store landed in a slot its reader has not emptied yet"
            ranks[dst].mail[parity][slot] = payload
        me = rng.randrange(2)
        r, peer = ranks[me], 1 - me
        if r.k >= n_exchanges:
            continue
        parity = r.k & 1
        if r.phase == "push":
            for slot in range(SLOTS):                       # 16-byte stores, each with its own flight time
                in_flight.append((now + rng.randrange(1, 30), peer, parity, slot, (r.k, (me, r.k, slot))))
            r.phase = "poll"
        elif r.phase == "poll":
            box = r.mail[parity]
            if all(v is not EMPTY for v in box):
                for slot, (k, value) in enumerate(box):
                    assert k == r.k and value == (peer, r.k, slot), "consumed a value of another exchange point"
                r.consumed.append(r.k)
                for slot in range(SLOTS):
                    box[slot] = EMPTY                       # emptied by its reader, within the same kernel
                r.k += 1
                r.phase = "push"
        max_drift = max(max_drift, abs(ranks[0].k - ranks[1].k))
    assert not in_flight and all(r.consumed == list(range(n_exchanges)) for r in ranks)
    return max_drift


@pytest.mark.parametrize("seed", range(25))
def test_mailbox_exchange_never_overwrites_and_never_reads_stale(seed):
    assert run(seed) <= 1      # a rank cannot pass exchange k before the peer has pushed k: two parities suffice


def test_one_parity_would_not_be_enough():
    # with a single mailbox per rank the schedule "A fast, B slow" collides: A consumes B's values of exchange 0, moves on and
    # pushes exchange 1 into B's mailbox while B has not yet consumed A's values of exchange 0
    mail_b = [EMPTY] * SLOTS
    for slot in range(SLOTS):
        mail_b[slot] = (0, ("A", 0, slot))       # A's push of exchange 0 has landed at B; B is slow and has not polled yet
    # B's push of exchange 0 landed at A, A consumed it and emptied its own mailbox: nothing stops A from pushing exchange 1
    collided = any(mail_b[slot] is not EMPTY for slot in range(SLOTS))
    assert collided
    # two parities: exchange 1 goes to the OTHER mailbox, and exchange 2 (same parity as 0) cannot be pushed before A has
    # consumed B's exchange 1, which B pushes only after it has emptied exchange 0 -- what the simulation above checks
