"""Model of the tensor regime's shared-memory ring protocol (tensor_regime.cu): one TMA producer, MMA issuers on
alternate tiles (or ONE issuer), mbarrier parity waits, nbuf accumulators drained in order by the epilogue.  The
agents are interleaved at random and TMA loads may LAND OUT OF ORDER (30 % of the completions pick a random in-flight
load, as loads to different DRAM channels / L2 hits and misses do).  sim() reports whether an issuer ever passes a
full-barrier wait on a stage that does not hold its tile's data -- the race behind the intermittent launch failures of
short rings.  With in-order landing (`ooo=0`) every geometry passes.

    python tools/ring_protocol_model.py        # table: geometry -> outcome over 200 random schedules

tests/test_ring_protocol.py checks the geometries rag::tensor::plan_ring() hands to the kernel against this model.
"""
import random


def parity_wait(completed_phases: int, parity: int) -> bool:
    """mbarrier.try_wait.parity: true once the phase with this parity has completed, i.e. the phase under way has
    the other parity (a fresh barrier passes parity 1)."""
    return (completed_phases & 1) != parity


def sim(n_ring: int, per_tile: int, n_tiles: int, nbuf: int, seed: int, one_issuer: bool = False, ooo: float = 0.3,
        straggler: bool = False) -> str:
    """straggler: now and then ONE in-flight load is held back for as long as anything else can make progress -- the
    worst case of out-of-order landing (a DRAM straggler), which random interleaving alone rarely produces."""
    rng = random.Random(seed)
    held = None                    # the in-flight load currently held back
    idle = 0                       # consecutive steps in which no agent made progress
    full_c = [0] * n_ring          # completed phases of the full / empty barrier of every stage
    empty_c = [0] * n_ring
    content = [None] * n_ring      # (tile, k-stage) the stage holds
    in_flight = []                 # TMA loads issued, not landed: (stage, (tile, k-stage))
    accf_c = [0] * nbuf
    acce_c = [0] * nbuf
    acc_content = [None] * nbuf
    prod = dict(t=0, ks=0, s=0, ph=0)
    issuers = [dict(it=0, s=0, ph=0, ks=0, state="start", id=i) for i in range(2)]
    mma_queue = []                 # MMAs issued, not retired: (issuer, what their commit signals, argument)
    epi_it = 0
    steps = 0
    sig = None
    while epi_it < n_tiles:
        now = (prod["t"], prod["ks"], len(in_flight), issuers[0]["it"], issuers[0]["ks"], issuers[0]["state"],
               issuers[1]["it"], issuers[1]["ks"], issuers[1]["state"], len(mma_queue), epi_it)
        if now != sig:
            idle = 0
            sig = now
        steps += 1
        if steps > 3_000_000:
            return "deadlock"
        agent = rng.choice(("prod", "tma", "i0", "i1", "mma", "epi"))
        idle += 1
        if straggler and held is None and in_flight and rng.random() < 0.002:
            held = rng.choice(in_flight)
        if agent == "tma" and held is not None:
            others = [x for x in in_flight if x is not held]
            if others:
                s, tag = others[rng.randrange(len(others)) if rng.random() < ooo else 0]
                in_flight.remove((s, tag))
                content[s] = tag
                full_c[s] += 1
                idle = 0
            elif idle > 400:       # nothing else has moved for a long time: the straggler finally lands
                in_flight.remove(held)
                content[held[0]] = held[1]
                full_c[held[0]] += 1
                held = None
                idle = 0
            continue
        if agent == "prod" and prod["t"] < n_tiles:
            s = prod["s"]
            if parity_wait(empty_c[s], prod["ph"] ^ 1):
                in_flight.append((s, (prod["t"], prod["ks"])))
                prod["s"] += 1
                if prod["s"] == n_ring:
                    prod["s"] = 0
                    prod["ph"] ^= 1
                prod["ks"] += 1
                if prod["ks"] == per_tile:
                    prod["ks"] = 0
                    prod["t"] += 1
        elif agent == "tma" and in_flight:
            j = rng.randrange(len(in_flight)) if rng.random() < ooo else 0
            s, tag = in_flight.pop(j)
            content[s] = tag
            full_c[s] += 1
        elif agent in ("i0", "i1"):
            x = issuers[int(agent[1])]
            if x["it"] >= n_tiles:
                continue
            it = x["it"]
            buf = it % nbuf
            par = (it // nbuf) & 1
            mine = (x["id"] == 0) if one_issuer else ((it & 1) == x["id"])
            if x["state"] == "start":
                if not mine:                       # the other issuer's tile: step over its stages, unseen
                    x["s"] += per_tile
                    if x["s"] >= n_ring:
                        x["s"] -= n_ring
                        x["ph"] ^= 1
                    x["it"] += 1
                    continue
                if parity_wait(acce_c[buf], par ^ 1):
                    x["state"] = "stages"
                    x["ks"] = 0
            else:
                s = x["s"]
                if parity_wait(full_c[s], x["ph"]):
                    if content[s] != (it, x["ks"]):
                        return (f"issuer {x['id']} passed the wait for tile {it} k-stage {x['ks']} on stage {s}, which holds "
                                f"{content[s]} (phases completed {full_c[s]}, parity waited {x['ph']})")
                    mma_queue.append((x["id"], "empty", s))
                    if x["ks"] == per_tile - 1:
                        mma_queue.append((x["id"], "accf", (buf, it)))
                    x["s"] += 1
                    if x["s"] == n_ring:
                        x["s"] = 0
                        x["ph"] ^= 1
                    x["ks"] += 1
                    if x["ks"] == per_tile:
                        x["state"] = "start"
                        x["it"] += 1
        elif agent == "mma" and mma_queue:          # the oldest MMA group of a random issuer retires: its commit arrives
            who = rng.choice([m[0] for m in mma_queue])
            k = next(i for i, m in enumerate(mma_queue) if m[0] == who)
            _, what, arg = mma_queue.pop(k)
            if what == "empty":
                empty_c[arg] += 1
            else:
                buf, it = arg
                accf_c[buf] += 1
                acc_content[buf] = it
        elif agent == "epi":
            buf = epi_it % nbuf
            par = (epi_it // nbuf) & 1
            if parity_wait(accf_c[buf], par):
                if acc_content[buf] != epi_it:
                    return f"epilogue drained tile {acc_content[buf]} where tile {epi_it} was due"
                acce_c[buf] += 1
                epi_it += 1
    return "ok"


def outcomes(n_ring, per_tile, nbuf, one_issuer=False, seeds=200, tiles=60, ooo=0.3, straggler=False):
    return {sim(n_ring, per_tile, tiles, nbuf, seed, one_issuer, ooo, straggler)[:40] for seed in range(seeds)}


if __name__ == "__main__":
    for nbuf in (2, 4):
        for per_tile in (2, 3):
            for n_ring in range(per_tile + 1, 15):
                two = outcomes(n_ring, per_tile, nbuf) | outcomes(n_ring, per_tile, nbuf, straggler=True)
                one = outcomes(n_ring, per_tile, nbuf, one_issuer=True, seeds=50) | outcomes(n_ring, per_tile, nbuf, one_issuer=True, seeds=50, straggler=True)
                print(f"stages {n_ring:2d}  per tile {per_tile}  accumulators {nbuf}:  two issuers "
                      f"{'ok' if two == {'ok'} else 'RACE'}   one issuer {'ok' if one == {'ok'} else 'RACE'}")
