"""CPU, build container only: the C oracle against the UNMODIFIED reference run headless
(oracle/ref_harness.py).  Skipped where /root/reference is absent (the GPU box)."""
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import common
from sugarscape_utilicalc_b200 import config_from_settings, synthetic

pytestmark = pytest.mark.reference
sys.path.insert(0, os.path.join(common.ROOT, "oracle"))


def _compare(s, rng, world, steps):
    import ref_harness as rh
    r = rh.ReferenceRun(s, rng=rng, world=world)
    s0 = r.snapshot()
    o = common.oracle_world(config_from_settings(s, rng_mode=rng))
    o.upload_cells(s0["sugar"], s0["cap"], s0["pollution"])
    o.upload_agents(s0["order"], s0["x"], s0["y"], s0["vision"], s0["metab"], s0["sugar_agent"])
    if rng == "mt":
        o.set_rng_mt19937(s0["mt"], s0["mt_index"])
    for t in range(steps):
        r.step()
        st_ = o.step(1)[0]
        rs = r.stats()
        assert st_["population"] == rs["population"][-1] and st_["gini"] == rs["gini"][-1]
        assert st_["metabolism_mean"] == rs["metabolismMean"][-1] and st_["vision_mean"] == rs["visionMean"][-1]
    sn, cells, ags = r.snapshot(), o.download_cells(), o.download_agents()
    exp = {k: sn[k] for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent")}
    common.assert_state_equal(cells, ags, exp, ordered=True, what="reference vs oracle")
    if rng == "mt":
        mt, idx = o.get_rng_mt19937()
        assert idx == sn["mt_index"] and np.array_equal(mt, sn["mt"])


def test_live_default_preset_mt():
    import ref_harness as rh
    _compare(rh.load_settings(), "mt", None, 12)


def test_live_pollution_preset_from_t0_keyed():
    import ref_harness as rh
    s = rh.load_settings("pollution")
    s["rule_params"]["pollution"]["pollution_start_time"] = 0
    s["rule_params"]["pollution"]["diffusion_start_time"] = 2
    _compare(s, "keyed", None, 15)


def test_utilicalc_spice_preset_crashes_in_the_reference():
    """parity for this preset is unpinned; the engine refuses it."""
    import ref_harness as rh
    s = rh.load_settings("utilicalc_spice")
    r = rh.ReferenceRun(s)
    with pytest.raises(KeyError):
        r.view.updateGame()
    from sugarscape_utilicalc_b200 import UnsupportedRuleError
    with pytest.raises(UnsupportedRuleError):
        config_from_settings(s)


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(n=st.integers(16, 40), seed=st.integers(0, 10 ** 6), pollution=st.booleans(), utilicalc=st.booleans(),
       seasons=st.booleans(), rng=st.sampled_from(["mt", "keyed"]), dens=st.sampled_from([0.05, 0.15, 0.4]))
def test_random_small_worlds(n, seed, pollution, utilicalc, seasons, rng, dens):
    if utilicalc:
        pollution = False
    w = synthetic.tiled_world(n, density=dens, seed=seed, vision=(1, min(6, (n - 1) // 2)))
    s = synthetic.synthetic_settings(n, pollution=pollution, seasons=seasons, utilicalc=utilicalc, seed=seed % 1000)
    if seasons:
        s["rule_params"]["seasons"]["period"] = 2
    world = dict(cap=w["cap"], sugar=w["sugar"], pollution=w["pollution"], x=w["x"], y=w["y"], vision=w["vision"],
                 metab=w["metab"], sugar_agent=w["sugar_agent"])
    _compare(s, rng, world, 6)


def test_dropin_install_glue_with_reference_view():
    """`dropin.install` on a live reference View (oracle injected as backend: no GPU here):
    stepping through the installed updateGame and syncing back equals the reference run."""
    import ref_harness as rh
    from sugarscape_utilicalc_b200 import dropin
    s = rh.load_settings("pollution")
    s["rule_params"]["pollution"]["pollution_start_time"] = 0
    ref = rh.ReferenceRun(s)
    ref.run(9)
    exp = ref.snapshot()
    exp_stats = ref.stats()
    run = rh.ReferenceRun(s)            # re-imports the reference: same seed, same initial state and MT stream
    h = dropin.install(run.ss, run.view, rng="mt", world_factory=common.oracle_world)
    for _ in range(9):
        run.view.updateGame()
    h.sync_back()
    got = run.snapshot()
    for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent", "mt"):
        assert np.array_equal(got[k], exp[k]), k
    assert got["mt_index"] == exp["mt_index"] and got["iteration"] == 9 and got["time"] == 9
    assert run.stats() == exp_stats
    # hand the View back: the reference's own loop continues from the synced state and RNG stream
    import random
    state9 = random.getstate()
    run.view.updateGame = h.original_update
    run.step()
    random.setstate(state9)                          # both runs share the one global `random`
    ref.step()
    a, b = run.snapshot(), ref.snapshot()
    for k in ("sugar", "pollution", "occ", "order", "x", "y", "sugar_agent", "mt"):
        assert np.array_equal(a[k], b[k]), k
