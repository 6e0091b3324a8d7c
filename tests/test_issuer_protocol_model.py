"""A small executable model of the barrier protocol inside tc_gemm_kernel / stem_pool_kernel (ncnn_b200/csrc/cuda/tc_gemm.cuh,
stem_pool.cuh): one TMA producer, one or two MMA-issuing warps on alternate tiles, an in-order epilogue (or two epilogue halves on
alternate tiles, the 32-wide-tile mode), a ring of S operand stages and A accumulator stages, every hand-over through an mbarrier
that is waited on BY PARITY -- the property that makes several issuers delicate: try_wait.parity(p) only says "the current phase has
parity != p", so a waiter that is two phases ahead of (or behind) its barrier is answered wrongly.

The model runs the agents under random interleavings with TMA loads landing out of order and checks, at every wait that passes, that
the thing waited for really happened (no stale pass) and that the run does not stop early (no deadlock).  It holds the rules the
kernels use to the claim DESIGN.md makes for them, and shows that the checker has teeth: the two variants that deadlocked on the GPU
(round 2) fail here as well.

  single    one issuer                                                                   -- must hold
  loose     two issuers whenever a tile takes <= half of the ring, no further sync       -- stale pass (MobileNetV2 144 -> 24 at full batch)
  observe   loose + each issuer parity-waits on the other's stages as it skips them      -- deadlock / stale (default bench, 3-stream leg)
  strict    two issuers only if a stage is reused >= A tiles later (stem_pool_kernel)    -- must hold
  words     loose + a monotonic "tiles observed" word per issuer (tc_gemm_kernel)        -- must hold

Pure Python, no GPU, no product code: this is a test of the DESIGN, kept next to the parity tests.
"""
import random

import pytest


class Stale(Exception):
    pass


class Bar(object):
    """an mbarrier reduced to what matters here: the index of its current (incomplete) phase"""

    def __init__(self):
        self.c = 0

    def passes(self, parity):
        return (self.c & 1) != parity


def simulate(variant, S, A, spt, tiles, halves, seed, max_steps=200000, issuers=2):
    rng = random.Random(seed)
    full = [Bar() for _ in range(S)]
    empty = [Bar() for _ in range(S)]
    tfull = [Bar() for _ in range(A)]
    tempty = [Bar() for _ in range(A)]
    loads_in_flight = []   # stage indices; any of them may land next (TMA completes out of order across stages)
    mma_fifo = []          # commit lists in issue order: the tensor pipe retires MMAs in the order they were queued
    obs = [-1] * issuers   # the monotonic words of the `words` variant

    if variant == "single":
        dual = False
    elif variant == "strict":
        dual = S >= A * spt and (not halves or S % (2 * spt) == 0)
    else:
        dual = 2 * spt <= S

    def producer():
        stage, phase = 0, 0
        for t in range(tiles):
            for j in range(spt):
                g = t * spt + j
                while not empty[stage].passes(phase ^ 1):
                    yield
                if empty[stage].c < g // S:
                    raise Stale("producer overwrote stage %d before use %d was released" % (stage, g // S - 1))
                loads_in_flight.append(stage)
                stage += 1
                if stage == S:
                    stage, phase = 0, phase ^ 1
                yield

    def issuer(me):
        stage, phase, acc, acc_phase = 0, 0, 0, 0
        for t in range(tiles):
            mine = ((t % issuers) == me) if dual else (me == 0)
            if not mine:
                for j in range(spt):
                    if variant == "observe" and dual:
                        while not full[stage].passes(phase):
                            yield
                    stage += 1
                    if stage == S:
                        stage, phase = 0, phase ^ 1
                acc += 1
                if acc == A:
                    acc, acc_phase = 0, acc_phase ^ 1
                continue
            while not tempty[acc].passes(acc_phase ^ 1):
                yield
            if tempty[acc].c < t // A:
                raise Stale("issuer %d reused accumulator stage %d of tile %d before the epilogue released it" % (me, acc, t))
            if variant == "words" and dual:
                for other in range(issuers):  # each other issuer's last tile before t
                    last = t - 1 - ((t - 1 - other) % issuers) if t > 0 else -1
                    while other != me and last >= 0 and obs[other] < last:
                        yield
            commits = []
            for j in range(spt):
                g = t * spt + j
                while not full[stage].passes(phase):
                    yield
                if full[stage].c < g // S + 1:
                    raise Stale("issuer %d passed the wait for stage %d use %d (tile %d) before it landed" % (me, stage, g // S, t))
                if variant == "words" and dual and j == spt - 1:
                    obs[me] = t
                commits.append(("empty", stage))
                stage += 1
                if stage == S:
                    stage, phase = 0, phase ^ 1
                yield
            commits.append(("tfull", acc))
            mma_fifo.append(commits)
            acc += 1
            if acc == A:
                acc, acc_phase = 0, acc_phase ^ 1
            yield

    done_tiles = [0]

    def epilogue(half):
        acc, acc_phase = 0, 0
        for t in range(tiles):
            if not halves or (t & 1) == half:
                while not tfull[acc].passes(acc_phase):
                    yield
                if tfull[acc].c < t // A + 1:
                    raise Stale("epilogue read accumulator stage %d of tile %d before its MMAs retired" % (acc, t))
                yield
                tempty[acc].c += 1
                done_tiles[0] += 1
            acc += 1
            if acc == A:
                acc, acc_phase = 0, acc_phase ^ 1

    agents = [producer()] + [issuer(i) for i in range(issuers)] + [epilogue(0)] + ([epilogue(1)] if halves else [])
    live = list(agents)
    idle = 0
    for _ in range(max_steps):
        choices = len(live) + (1 if loads_in_flight else 0) + (1 if mma_fifo else 0)
        if done_tiles[0] == tiles:
            return "ok"
        before = (tuple(b.c for b in full + empty + tfull + tempty), len(loads_in_flight), len(mma_fifo), tuple(obs), done_tiles[0])
        k = rng.randrange(choices)
        if k < len(live):
            try:
                next(live[k])
            except StopIteration:
                live.pop(k)
        elif k == len(live) and loads_in_flight:
            s = loads_in_flight.pop(rng.randrange(len(loads_in_flight)))
            full[s].c += 1
        else:
            for kind, i in mma_fifo.pop(0):
                (empty if kind == "empty" else tfull)[i].c += 1
        after = (tuple(b.c for b in full + empty + tfull + tempty), len(loads_in_flight), len(mma_fifo), tuple(obs), done_tiles[0])
        # an agent that only spun leaves the state unchanged; nothing in flight and nobody moving for long = deadlock
        idle = idle + 1 if (before == after and not loads_in_flight and not mma_fifo) else 0
        if idle > 400:
            return "deadlock"
    return "deadlock"


# (ring stages, accumulator stages, stages per tile, alternate epilogue halves): shapes the kernels really run --
# 64-wide shifted-window tiles (ring of 3-4 big stages, 8 accumulator stages, 1-2 slabs), 256-wide tiles (2 accumulator stages),
# 128-wide (4), the stem (ring >= 8, 8 accumulator stages, 1 stage per tile), 32-wide tiles read by alternate halves
CONFIGS = [(3, 8, 1, False), (4, 8, 1, False), (4, 8, 2, False), (8, 8, 1, False), (16, 8, 1, False), (16, 8, 2, False), (16, 4, 4, False),
           (16, 4, 8, False), (16, 2, 8, False), (6, 2, 3, False), (5, 4, 2, False), (16, 8, 1, True), (8, 8, 1, True), (16, 8, 2, True), (12, 8, 2, True), (6, 8, 3, True), (16, 8, 9, False)]


def outcomes(variant, seeds):
    res = {}
    for cfg in CONFIGS:
        S, A, spt, halves = cfg
        for seed in seeds:
            try:
                r = simulate(variant, S, A, spt, tiles=40, halves=halves, seed=seed)
            except Stale as e:
                r = "stale: %s" % e
            if r != "ok":
                res.setdefault(cfg, r)
    return res


@pytest.mark.parametrize("variant", ["single", "strict", "words"])
def test_protocols_in_the_tree_hold(variant):
    bad = outcomes(variant, range(60))
    assert not bad, bad


def test_checker_catches_the_variants_that_failed_on_the_gpu():
    loose = outcomes("loose", range(60))
    assert any(r.startswith("stale") for r in loose.values()), loose          # the MobileNetV2 full-batch failure
    observe = outcomes("observe", range(60))
    assert observe, observe                                                    # the default-bench failure (deadlock or stale pass)
    assert any(r == "deadlock" for r in observe.values()) or any(r.startswith("stale") for r in observe.values())


@pytest.mark.parametrize("variant", ["strict", "words"])
def test_three_issuers_of_the_fused_stem(variant):
    """stem_pool_kernel: three issuers, 8 accumulator stages (an accumulator stage comes back to a DIFFERENT issuer: 8 % 3 != 0),
    one ring stage per conv row, ring of 8 or more"""
    for S in (8, 10, 16):
        for seed in range(40):
            assert simulate(variant, S, 8, 1, tiles=50, halves=False, seed=seed, issuers=3) == "ok", (S, seed)


def test_words_protocol_uses_two_issuers_where_strict_cannot():
    """the point of the monotonic words: the 64-wide shifted-window shapes get their second issuer back"""
    for S, A, spt in [(3, 8, 1), (4, 8, 1), (4, 8, 2)]:
        assert 2 * spt <= S and not (S >= A * spt)
