"""The plain-C oracle against the committed golden vectors (tests/golden/*.npz, generated from the unmodified
reference by tests/golden/make_golden.py).  Runs anywhere: needs neither /root/reference nor a GPU."""
import numpy as np
import pytest

import cpu_harness as H
import golden_io as G
import scenarios as S


@pytest.mark.parametrize("name", G.names())
def test_oracle_reproduces_reference_golden(name):
    g = G.load(name)
    sc = S.by_name(name)
    assert S.params_json(sc.params) == S.params_json(g["params"]), "scenario definition drifted from the fixture: regenerate"
    assert len(sc.steps) == g["n_steps"]
    O = H.oracle(fresh=True)
    O.replay(g["all_draws"])
    e = S.new_emitter(O, sc)
    snaps, _ = S.run(e, sc)
    assert O.rand_replay_remaining() == 0, "the oracle consumed a different number of draws than the reference"
    S.compare(snaps, g["snaps"], exact=True, what=name)  # CPU vs CPU: bit-exact everywhere, trig included


def test_golden_covers_the_edge_cases():
    names = set(G.names())
    assert {"fire", "smoke", "explosion", "burst_extinction", "capacity_saturation", "axis_rotation",
            "stop_and_drain", "small_dt_rate_cap"} <= names
    g = G.load("stop_and_drain")
    assert g["snaps"][-1]["count"] == 0 and not g["snaps"][-1]["active"]
    g = G.load("capacity_saturation")
    assert g["snaps"][-1]["count"] == g["params"].particle_count_max
    g = G.load("burst_extinction")
    counts = [s["count"] for s in g["snaps"]]
    assert counts[0] == g["params"].particle_count_max and counts[-1] < 50 and sorted(counts, reverse=True) == counts
