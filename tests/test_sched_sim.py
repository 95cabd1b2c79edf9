"""tools/sched_sim.py (discrete-event model of the staged scheduler): the dependency protocol it shares with the kernel
must complete every macroblock in every variant, including the fused-deblocking protocol kept in tools/experiments."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("variant", ["base", "patience", "dfull", "dfill", "half", "fastB"])
def test_every_variant_completes(variant):
    import sched_sim
    r = sched_sim.simulate(n_streams=6, warps=24, n_cta=8, variant=variant)      # asserts completion internally
    assert r["kernel_ms"] > 0 and 0 < r["warp_busy_frac"] <= 1.0


def test_more_streams_raise_utilisation():
    import sched_sim
    a = sched_sim.simulate(n_streams=4, warps=24, n_cta=16)
    b = sched_sim.simulate(n_streams=16, warps=24, n_cta=16)
    assert b["warp_busy_frac"] > a["warp_busy_frac"]
