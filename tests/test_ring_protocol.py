"""Host logic of the tensor regime's shared-memory ring (rag::tensor::plan_ring through rag_debug_ring_plan; no device
needed) against the protocol model in tools/ring_protocol_model.py: every geometry the planner hands to the kernel
must survive random schedules with out-of-order TMA completion, and the model must still catch the geometries that
raced on the GPU (4-5 stages with 3 per tile)."""
import ctypes as C
import importlib.util
import os

import pytest

from local_rag_system_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ring_protocol_model", os.path.join(ROOT, "tools", "ring_protocol_model.py"))
model = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(model)


def plan(avail, per_tile, nbuf):
    used, one = C.c_int32(), C.c_int32()
    _native.check(_native.load().rag_debug_ring_plan(avail, per_tile, nbuf, C.byref(used), C.byref(one)))
    return used.value, bool(one.value)


@pytest.mark.parametrize("per_tile", [1, 2, 3])
@pytest.mark.parametrize("nbuf", [2, 4])
def test_planned_rings_survive_the_protocol_model(per_tile, nbuf):
    for avail in range(per_tile + 1, 15):
        used, one = plan(avail, per_tile, nbuf)
        assert per_tile + 1 <= used <= avail
        assert one or used >= 2 * per_tile                      # two issuers never share a ring shorter than two tiles
        got = model.outcomes(used, per_tile, nbuf, one_issuer=one, seeds=40, tiles=40)
        got |= model.outcomes(used, per_tile, nbuf, one_issuer=one, seeds=60, tiles=40, straggler=True)
        assert got == {"ok"}, (avail, per_tile, nbuf, used, one, got)


def test_planner_keeps_the_production_geometries():
    assert plan(14, 3, 2) == (14, False)        # cta_group::2 pairs at D = 768: the headline batch-1024 kernel
    assert plan(7, 3, 2) == (7, False)          # single CTAs at D = 768
    assert plan(14, 2, 4) == (14, False)        # pairs at D = 384
    assert plan(7, 2, 4) == (4, False)          # single CTAs at D = 384: 7 stages could alias against tile t-3 -> two whole tiles
    assert plan(6, 2, 4) == (4, False)          # k <= 128 heaps next to the ring (config 5)
    assert plan(5, 3, 2) == (5, True)           # fp32 filter buffers at D = 768, single CTAs: one issuer
    assert plan(4, 3, 2) == (4, True)


@pytest.mark.parametrize("n_ring,per_tile,nbuf", [(3, 2, 2), (4, 3, 2), (5, 3, 2), (5, 2, 4)])
def test_model_catches_the_rings_that_raced(n_ring, per_tile, nbuf):
    got = model.outcomes(n_ring, per_tile, nbuf, seeds=200, tiles=60)
    assert got != {"ok"} and "deadlock" not in got
    assert model.outcomes(n_ring, per_tile, nbuf, seeds=50, tiles=60, ooo=0.0) == {"ok"}      # in-order landing hides it
    assert model.outcomes(n_ring, per_tile, nbuf, one_issuer=True, seeds=50, tiles=60) == {"ok"}


def test_planner_is_not_more_cautious_than_the_model():
    """Every geometry the planner refuses to run with two issuers does race in the model once one load is held back
    (a DRAM straggler) -- including round 1's default of 7 stages x 2 per tile x 4 accumulators."""
    for per_tile, nbuf in ((2, 2), (3, 2), (2, 4), (3, 4)):
        for avail in range(per_tile + 1, 15):
            used, one = plan(avail, per_tile, nbuf)
            if used == avail and not one:
                continue                                       # taken as is: covered by the test above
            got = model.outcomes(avail, per_tile, nbuf, seeds=150, tiles=60, straggler=True)
            assert got != {"ok"}, (avail, per_tile, nbuf, "the model finds no race with two issuers here")


def test_ring_plan_rejects_nonsense():
    with pytest.raises(ValueError):
        plan(2, 3, 2)
    with pytest.raises(ValueError):
        plan(7, 2, 3)
