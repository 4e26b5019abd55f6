"""Pins the CPU oracle to the reference: replays every committed golden fixture (outputs of the
live Python reference, tests/golden/make_golden.py) through oracle/inversus_oracle.c.

Integer state, obs, done and info flags must be bit-exact; rewards are checked both to the
north-star tolerance (1e-6 relative) and -- stronger -- for exact float32 equality.
"""
import os

import numpy as np
import pytest

from backends import OracleBackend
from golden.scenarios import SCENARIOS, compare, run_scenario

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_reproduces_reference_fixture(name):
    sc = SCENARIOS[name]
    gold = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    rec = run_scenario(OracleBackend(sc), sc)
    compare(rec, gold, float_rtol=1e-6, what=name)
    # the fp64 accumulation order is restated exactly: the float32 rewards (compare() above) and
    # the running fp64 episode returns match bit for bit, not just to 1e-6
    assert np.array_equal(rec["reward_f32"], gold["reward_f32"])
    assert np.array_equal(rec["episode_return"], gold["episode_return"])


def test_fixtures_cover_the_edge_cases():
    """The fixtures must actually contain the situations the parity claim is about."""
    g = {n: np.load(os.path.join(GOLD, f"{n}.npz")) for n in SCENARIOS}
    flags = np.concatenate([g[n]["flags"].ravel() for n in g])
    assert (flags & 1).any() and (flags & 2).any() and (flags & 4).any() and (flags & 8).any()
    # ties: done by round-over with neither win nor lose, and timeouts
    hs = g["hard_short"]
    timeouts = (hs["done"] == 1) & (hs["episode_steps"] == 40) & ((hs["flags"] & 12) == 0)
    assert timeouts.any()
    # charge shots fired, >= 9 simultaneous bullets, both-dead ties somewhere
    assert g["hard_charge"]["state"]["n_bullets"].max() >= 9
    assert g["selfplay_shooty"]["state"]["n_bullets"].max() >= 10
    # table scenario: P2 spawned on top of P1 after 20 failed tries
    st = g["hard_table"]["init_state"]
    assert ((st["p1"][:, 0] == st["p2"][:, 0]) & (st["p1"][:, 1] == st["p2"][:, 1])).any()
    # stepping past done without reset (manual-reset scenario)
    hm = g["hard_manual"]
    assert (hm["episode_steps"] > 60).any()
