"""Model check of the one-sided exchange's slot / flag protocol (engine.py, csrc/cf_p2p.cu).

No GPU and no library call: a randomised discrete-event simulation of W ranks running T steps of L layers,
with the dependency structure the engine builds --

  serial step     (PatchGatherEngine.step):            put(0) apply(0) put(1) apply(1) ...  on one stream
  two-chain step  (PatchGatherEngine._step_overlapped): put(l) on the main stream after apply(l - lag),
                                                        apply(l) on the side stream after put(l)
  ring order      (RingExchangeEngine.exchange):        put(l), then one apply per origin, own shard first

-- and the protocol's rules: put_t(l) of rank A rewrites slot (l, A) in EVERY rank's receive region and then
bumps that slot's flag; apply_t(l) of rank B may start once flag(l, A) >= B's own put count of layer l for
every origin A it consumes, and reads the slots while it runs; a step (one CUDA graph launch) starts after
the rank's previous step has finished.  Kernels are not atomic: a put and an apply are 'running' between a
start and an end event, and the scheduler interleaves all enabled events at random.

Checked: an apply of step t only ever sees payloads of step t (no slot is rewritten before or while the
peer that still needs it reads it), and every schedule runs to completion (no deadlock).  The same checker
must FIND the race when the compress chain is allowed to run arbitrarily far ahead, which is what the
engine's OVERLAP_LAG / `layers >= 2 * lag + 1` rule excludes.
"""
import random

import pytest


class Violation(Exception):
    pass


def template_for(mode, rank, world, layers, lag):
    """One step of one rank as a list of entries {kind, l, origins, stream, deps (indices of earlier entries)}.
    Stream order is implicit: an entry also depends on the previous entry of its stream."""
    tpl, applies = [], {}

    def add(kind, l, stream, deps, origins=None):
        tpl.append(dict(kind=kind, l=l, stream=stream, deps=list(deps), origins=origins))
        return len(tpl) - 1

    for l in range(layers):
        if mode == "serial":
            p = add("put", l, "main", [])
            applies[l] = add("apply", l, "main", [p], origins=tuple(range(world)))
        elif mode == "ring":
            p = add("put", l, "main", [])
            for hop in range(world):
                applies[l] = add("apply", l, "main", [p], origins=((rank - hop) % world,))
        else:  # two chains
            p = add("put", l, "main", [applies[l - lag]] if l >= lag else [])
            applies[l] = add("apply", l, "side", [p], origins=tuple(range(world)))
    return tpl


def instantiate(rank, template, steps):
    """Ops of one rank over `steps` graph launches: ids (rank, t, index); the first entry of every stream
    depends on the tail of every stream of the previous step (a graph launch starts after the previous one
    has finished)."""
    ops, prev_tail = [], []
    for t in range(steps):
        last_on = {}
        for idx, e in enumerate(template):
            deps = {(rank, t, d) for d in e["deps"]}
            deps |= {last_on[e["stream"]]} if e["stream"] in last_on else set(prev_tail)
            last_on[e["stream"]] = (rank, t, idx)
            ops.append(dict(id=(rank, t, idx), kind=e["kind"], t=t, l=e["l"], origins=e["origins"], deps=deps))
        prev_tail = list(last_on.values())
    return ops


def simulate(world, layers, steps, mode, lag, seed, templates=None):
    """`templates`: per-rank step templates (e.g. recorded from the engine); default: `template_for(mode, ...)`."""
    rng = random.Random(seed)
    ops = {}
    for r in range(world):
        tpl = templates[r] if templates is not None else template_for(mode, r, world, layers, lag)
        for op in instantiate(r, tpl, steps):
            ops[op["id"]] = op
    # slot[b][(l, a)] = step whose payload it holds (-1: the warm-up state); flag counts completed puts
    version = [{(l, a): -1 for l in range(layers) for a in range(world)} for _ in range(world)]
    writing = [{k: False for k in version[b]} for b in range(world)]
    readers = [{k: 0 for k in version[b]} for b in range(world)]
    flag = [{k: 0 for k in version[b]} for b in range(world)]
    puts_done = [[0] * layers for _ in range(world)]
    pending, running, ended = set(ops), set(), set()

    def startable(op):
        if not op["deps"] <= ended:
            return False
        if op["kind"] == "apply":  # device-side wait: flag >= own put count of this layer
            b = op["id"][0]
            return all(flag[b][(op["l"], a)] >= puts_done[b][op["l"]] for a in op["origins"])
        return True

    while pending or running:
        events = [("start", i) for i in pending if startable(ops[i])] + [("end", i) for i in running]
        if not events:
            raise Violation(f"deadlock with {len(pending)} ops pending")
        what, i = rng.choice(sorted(events))
        op = ops[i]
        a, t, l = i[0], op["t"], op["l"]
        if what == "start":
            pending.discard(i)
            running.add(i)
            if op["kind"] == "put":
                for b in range(world):
                    if readers[b][(l, a)]:
                        raise Violation(f"put {i} rewrites slot ({l},{a}) of rank {b} while it is being read")
                    writing[b][(l, a)] = True
                    version[b][(l, a)] = t  # bytes start landing: from now on the slot is no longer step t-1's
            else:
                for o in op["origins"]:
                    if writing[a][(l, o)] or version[a][(l, o)] != t:
                        raise Violation(f"apply {i} reads slot ({l},{o}) holding step {version[a][(l, o)]} "
                                        f"(writing={writing[a][(l, o)]})")
                    readers[a][(l, o)] += 1
        else:
            running.discard(i)
            ended.add(i)
            if op["kind"] == "put":
                for b in range(world):
                    writing[b][(l, a)] = False
                    flag[b][(l, a)] += 1
                puts_done[a][l] += 1
            else:
                for o in op["origins"]:
                    if version[a][(l, o)] != t:
                        raise Violation(f"slot ({l},{o}) of rank {a} was rewritten while apply {i} was reading it")
                    readers[a][(l, o)] -= 1
    return len(ended)


def _engine_rule():
    from compactfusion_b200.engine import PatchGatherEngine
    return PatchGatherEngine.OVERLAP_LAG


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("mode", ["serial", "ring"])
def test_single_chain_schedules_never_reuse_a_live_slot(world, mode):
    for layers in (2, 3, 5):  # the engine uses the one-sided transport from 2 layers up
        for seed in range(40):
            n = simulate(world, layers, steps=3, mode=mode, lag=0, seed=seed)
            assert n == world * 3 * layers * (1 + (world if mode == "ring" else 1))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_two_chain_schedule_is_safe_under_the_engines_rule(world):
    lag = _engine_rule()
    for layers in (2 * lag + 1, 2 * lag + 2, 9):
        for seed in range(60):
            simulate(world, layers, steps=3, mode="overlap", lag=lag, seed=seed)


def test_checker_finds_the_race_when_the_lag_is_unbounded():
    """Teeth: with the compress chain free to run a whole step ahead, some schedule lets a peer rewrite a slot
    that the slow rank has not reconstructed yet -- and fewer layers than 2 * lag + 1 break the lag rule too."""
    def finds(layers, lag, world=2, seeds=300):
        for seed in range(seeds):
            try:
                simulate(world, layers, steps=3, mode="overlap", lag=lag, seed=seed)
            except Violation as e:
                assert "deadlock" not in str(e)
                return True
        return False

    assert finds(layers=4, lag=4)   # lag >= layers: no put ever waits for a reconstruct
    lag = _engine_rule()
    assert finds(layers=2 * lag, lag=lag) or finds(layers=2 * lag - 1, lag=lag)


def test_single_layer_is_unsafe_which_is_why_the_engine_falls_back_to_nccl():
    """layers == 1: a fast rank's next put can land while the peer still reads (engine.py: transport stays
    NCCL below 2 layers)."""
    found = False
    for seed in range(300):
        try:
            simulate(2, 1, steps=3, mode="serial", lag=0, seed=seed)
        except Violation:
            found = True
            break
    assert found
