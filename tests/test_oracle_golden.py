"""CPU: the C oracle against the golden vectors the unmodified reference produced."""
import pytest

import common


@pytest.mark.parametrize("name", common.GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    g = common.load_golden(name)
    w = common.world_from_golden(g, "oracle")
    assert common.run_against_golden(g, w, exact_gini=True) == g["steps"]
    w.close()


@pytest.mark.parametrize("name", common.LIFE_CASES)
def test_oracle_lifespan_and_procreation_match_reference_golden(name):
    """examples/evolution_metabolism (= examples/tag_flipping) as shipped, and procreation without the
    lifespan rule -- cells, agents in list order (newborns act in the step they are born in, an agent that reached maxAge
    stays for one more turn), age / sex / fertility / endowment / tags, Environment.idIncrement, the event lines and the
    MT19937 state, step by step against the unmodified reference."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "oracle")
    assert common.run_against_golden(g, w, exact_gini=True, chunk=3) == g["steps"]
    w.close()


@pytest.mark.parametrize("name", common.SPICE_CASES)
def test_oracle_spice_matches_reference_golden(name):
    """examples/trade_price_volume and examples/readme_example without the trade rule -- spice on the cells and
    in the agents, the Cobb-Douglas welfare of move() (w1 ** (m1/mt) * w2 ** (m2/mt), compared for the arg-max: x ** y is the
    reference's libm pow bit for bit, csrc/ss_pow.h), spice metabolism and starvation, statistics -- step by step against
    the unmodified reference."""
    g = common.load_golden(name)
    w = common.world_from_golden(g, "oracle")
    assert common.run_against_golden(g, w, exact_gini=True, chunk=5) == g["steps"]
    assert w.pow_faults() == 0
    w.close()


def test_pow_restatement_equals_libm_and_python():
    """csrc/ss_pow.h (glibc 2.39 __pow_fma restated from its disassembly, tables read out of libm) against this machine's
    libm pow() -- which is what CPython's float ** float calls -- on random operands of the welfare domain."""
    import ctypes
    import math
    import random
    import subprocess
    import os
    d = os.path.join(common.ROOT, "oracle", "pow_glibc")
    exe = os.path.join(d, "test_pow")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(d, "test_pow.c"), "-lm"])
    out = subprocess.check_output([exe, "2000000"]).decode()
    assert out.strip().endswith(" 0 mismatches"), out
    rng = random.Random(5)
    for _ in range(2000):                                   # Python's own ** is that libm pow
        x, m1, m2 = rng.randint(1, 500) * rng.choice([1.0, 0.25]), rng.randint(1, 9), rng.randint(1, 9)
        assert x ** (m1 / (m1 + m2)) == math.pow(x, m1 / (m1 + m2))


def test_oracle_chunked_steps_equal_single_steps():
    g = common.load_golden("c2_pollution_mt")
    w = common.world_from_golden(g, "oracle")
    common.run_against_golden(g, w, exact_gini=True, chunk=7, max_steps=100)
    w.close()


@pytest.mark.parametrize("name", common.OUTPUT_CASES)
def test_run_recorder_outputs_match_reference_golden(name, tmp_path):
    common.check_outputs_against_golden(name, "oracle", tmp_path)


@pytest.mark.parametrize("name", ["c2_pollution_mt", "c2_seasons_keyed", "c1_default_utilicalc_mt", "f3_evolution_metabolism_mt", "f2_spice_life_mt"])
def test_checkpoint_resume_is_bit_identical(name, tmp_path):
    """checkpoint.save / restore (here over the CPU oracle): a run stopped at step 7 and resumed from the file ends
    where the uninterrupted run ends -- cells, agents in list order, statistics, MT19937 state."""
    common.check_checkpoint_resume(name, "oracle", tmp_path)


def test_checkpoint_resume_keeps_the_order_width_when_the_top_id_died(tmp_path):
    common.check_resume_after_top_id_died("oracle", tmp_path)


def test_run_recorder_started_on_a_resumed_run_has_no_history(tmp_path):
    """RunRecorder(iteration=250) must not index a population series it does not have."""
    from sugarscape_utilicalc_b200 import outputs
    g = common.load_golden("c1_default_moveeat_keyed")
    w = common.world_from_golden(g, "oracle")
    rec = outputs.RunRecorder(g["settings"], w, rng="keyed", data_file=str(tmp_path / "data.txt"), iteration=250)
    rec.step()
    assert rec.iteration == 251 and len(rec.population) == 1
