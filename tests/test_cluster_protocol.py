"""The tick protocol of the cluster column kernels (templates/template_column.cu, ColumnPlan._cluster), replayed on
a model of mbarriers: a producer issues ring slot e only after every peer has issued slot e - SLACK; one remote
arrive per peer and slot on tick[e % TICKS], waits on local barriers by phase parity only.  Checked under random
schedules: no deadlock, the drift bound, no wait that passes on the parity of a *later* phase (what 2 * SLACK + 1
barriers are for), no arrive on a CTA that has already left."""

import random

import pytest


class Barrier:
    """mbarrier with an arrival count: the phase completes (and its parity flips) when `count` arrivals are in."""

    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count

    def try_wait(self, parity):
        # completes when the phase with this parity is over: the current phase has the other parity
        return (self.phase & 1) != parity


def run(ranks, slack, ticks, slots, rng):
    tick = [[Barrier(ranks - 1) for _ in range(ticks)] for _ in range(ranks)]
    issued = [0] * ranks                      # slots issued so far per producer
    state = ["wait"] * ranks                  # wait -> (issue + arrives) per slot, then "drain", then "gone"
    drain = [0] * ranks
    steps = 0
    while any(s != "gone" for s in state):
        steps += 1
        assert steps < 200000, "deadlock: %s %s" % (state, issued)
        r = rng.randrange(ranks)
        if state[r] == "wait":
            e = issued[r]
            if e >= slack:
                d = e - slack
                if not tick[r][d % ticks].try_wait((d // ticks) & 1):
                    continue
                # the wait must not have passed on a stale or future phase: every peer has really issued slot d
                assert all(issued[p] > d for p in range(ranks) if p != r), (r, e, issued)
            assert all(issued[r] - issued[p] <= slack for p in range(ranks)), "drift bound broken"
            for p in range(ranks):
                if p != r:
                    assert state[p] != "gone", "arrive on a CTA that has exited"
                    tick[p][e % ticks].arrive()
            issued[r] += 1
            if issued[r] == slots:
                state[r], drain[r] = "drain", max(slots - slack, 0)
        elif state[r] == "drain":
            d = drain[r]
            if d == slots:
                state[r] = "gone"
            elif tick[r][d % ticks].try_wait((d // ticks) & 1):
                assert all(issued[p] > d for p in range(ranks) if p != r)
                drain[r] += 1
    return steps


@pytest.mark.parametrize("ranks,slack", [(2, 1), (2, 2), (4, 1), (4, 4), (8, 2), (3, 3)])
def test_tick_protocol_is_deadlock_free_and_bounded(ranks, slack):
    rng = random.Random(ranks * 100 + slack)
    for slots in (1, slack, 2 * slack + 1, 7 * (2 * slack + 1) + 3):
        for _ in range(20):
            run(ranks, slack, 2 * slack + 1, slots, rng)


def test_fewer_tick_barriers_would_alias():
    """With only 2 * SLACK barriers the check fires under some schedule (a wait passes on the parity of a phase that
    is two rounds old / the barrier is reused while a slower peer still waits on it) -- or the model deadlocks."""
    rng = random.Random(7)
    with pytest.raises(AssertionError):
        for _ in range(300):
            run(3, 2, 2 * 2 - 1, 60, rng)


def test_plan_allocates_the_barriers_the_protocol_needs():
    from absinthe_b200 import programs as P
    from absinthe_b200.emitter import Plan
    f = P.KERNELS["fastwaves"][1]()
    q = P.make_program("fastwaves", [f], [(2, 4, 6)], domain=(128, 64, 60))
    q["CUDA"] = {"STREAM_K": "regs", "COLUMN": {"WARPS": (2, 3), "MERGE": True, "CLUSTER": (2, 1), "SLACK": 3}}
    ctx = Plan(q).groups[0].context()
    assert ctx["TICKS"] == 2 * ctx["SLACK"] + 1 == 7 and ctx["CLUSTER"] == 2
    assert ctx["TICK_OFFSET"] >= ctx["BARRIER_OFFSET"] + 16 * ctx["NS"]
    assert ctx["TICK_OFFSET"] + 8 * ctx["TICKS"] <= ctx["SMEM_BYTES"]
