"""Pins the pure-Python restatement of the reference's step loop (oracle/py_loop.py) to the
reference: it must reproduce the committed golden fixtures -- outputs of the live Python reference
-- bit for bit (state, ordered bullets, observations, flags, fp64 rewards and returns)."""
import os

import numpy as np
import pytest

from backends import PyLoopBackend
from golden.scenarios import SCENARIOS, compare, run_scenario

GOLD = os.path.join(os.path.dirname(__file__), "golden")
# the long scenarios are covered by the C oracle; the Python loop replays the shorter ones in full
NAMES = ["hard_short", "hard_manual", "hard_table", "easy_table", "hard_charge", "selfplay_shooty"]


@pytest.mark.parametrize("name", NAMES)
def test_python_loop_reproduces_reference_fixture(name):
    sc = SCENARIOS[name]
    gold = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    rec = run_scenario(PyLoopBackend(sc), sc)
    compare(rec, gold, float_rtol=1e-6, what=name)
    assert np.array_equal(rec["reward"], gold["reward"])              # fp64 rewards, exactly
    assert np.array_equal(rec["episode_return"], gold["episode_return"])
