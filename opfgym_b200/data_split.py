"""Train / validation / test split of the 35 136 quarter-hour SimBench steps.

Host-side index sets only (reference ``opfgym/simbench/data_split.py:5-59``):
deterministic weekly blocks, equidistant over the 52 weeks, first for the test
set and then -- among the remaining weeks -- for the validation set.
"""
from __future__ import annotations

import numpy as np

N_STEPS = 24 * 4 * 366
WEEK = 7 * 24 * 4


def _week_blocks(week_ids) -> np.ndarray:
    if len(week_ids) == 0:
        return np.array([], dtype=np.int64)
    return np.concatenate([np.arange(w * WEEK, (w + 1) * WEEK) for w in week_ids])


def define_test_train_split(test_share=0.2, random_test_steps=False, validation_share=0.2,
                            random_validation_steps=False, rng=None, **_):
    """Returns ``(test_steps, validation_steps, train_steps)``."""
    assert test_share + validation_share <= 1.0
    if random_test_steps:
        assert random_validation_steps, "random test data needs random validation data"
    rng = rng or np.random
    steps = np.arange(N_STEPS)
    empty = np.array([], dtype=np.int64)
    if test_share == 1.0:
        return steps, empty, empty
    test_weeks = np.array([], dtype=int)
    if test_share == 0.0:
        test = empty
    elif random_test_steps:
        test = rng.choice(steps, int(N_STEPS * test_share))
    else:
        test_weeks = np.linspace(0, 51, num=int(52 * test_share), dtype=int)
        test = _week_blocks(test_weeks)
    remaining = np.setdiff1d(steps, test)
    if validation_share == 1.0:
        return empty, steps, empty
    if validation_share == 0.0:
        val = empty
    elif random_validation_steps:
        val = rng.choice(remaining, int(N_STEPS * validation_share))
    else:
        free_weeks = np.setdiff1d(np.arange(52), test_weeks)
        pick = np.linspace(0, len(free_weeks) - 1, num=int(52 * validation_share), dtype=int)
        val = _week_blocks(free_weeks[pick])
    train = np.setdiff1d(remaining, val)
    return test, val, train
