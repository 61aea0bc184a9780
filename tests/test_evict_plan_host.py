"""Host side of area 2 (no GPU): kvq_evict_plan_make / kvq_evict_indices_host against the golden index
sets of the reference and against the oracle on a random sweep; error behaviour of the trim_* API."""

import random

import numpy as np
import pytest

from _golden import load
from kvq_b200 import _lib as L
from kvq_b200.cache import evict_plan, kept_token_indices


def ours(st, T, w, p, s, blk, keep, bud):
    return kept_token_indices(evict_plan(st, T, window=w, prefix=p, stride=s, block=blk, keep=keep, budget=bud))


def test_golden_index_sets_bit_exact():
    g = load("evict.npz")
    params, offs, flat = g["params"], g["offsets"], g["indices"]
    for i, row in enumerate(params.tolist()):
        want = flat[offs[i]:offs[i + 1]].tolist()
        assert ours(*row) == want, row


def test_random_sweep_against_oracle(oracle):
    rng = random.Random(7)
    for _ in range(3000):
        T = rng.choice([rng.randint(1, 60), rng.randint(1, 3000), rng.randint(1, 200000)])
        w = rng.choice([0, 1, rng.randint(0, 64), rng.randint(0, 600)])
        p = rng.choice([0, rng.randint(0, 8), rng.randint(0, 100)])
        st = rng.randint(0, 4)
        s = rng.randint(1, 9)
        blk = rng.randint(1, 130)
        keep = rng.randint(1, blk)
        bud = rng.choice([0, 1, 2, rng.randint(0, 100), rng.randint(0, max(1, T))])
        want = oracle.evict_indices(st, T, window=w, prefix=p, stride=s, block=blk, keep=keep, budget=bud).tolist()
        assert ours(st, T, w, p, s, blk, keep, bud) == want, (st, T, w, p, s, blk, keep, bud)


def test_plan_counts_at_long_context():
    # C4 kept counts at T = 131072, defaults of benchmarker.py:648-656
    assert evict_plan(L.STRIDED, 131072, window=256, stride=4).n_total == 32960
    assert evict_plan(L.BLOCK, 131072, window=256, block=64, keep=8).n_total == 16608
    assert evict_plan(L.BUDGET, 131072, window=256, budget=64).n_total == 320
    assert evict_plan(L.SLIDING, 131072, window=256).n_total == 256


def test_duplicates_need_a_table(oracle):
    """At ~1M tokens with a budget just under the older region, float32 linspace(...).long() repeats
    values and unique_consecutive drops them: the plan must count them out and ask for table mode.
    Golden: the reference itself on those parameters (evict_big.npz, delta encoded)."""
    g = load("evict_big.npz")
    saw_table = False
    for i, (T, w, p, bud) in enumerate(g["params"].tolist()):
        want = np.concatenate([[int(g[f"{i}/first"])], int(g[f"{i}/first"]) + np.cumsum(g[f"{i}/deltas"].astype(np.int64))])
        plan = evict_plan(L.BUDGET, T, window=w, prefix=p, budget=bud)
        got = np.array(kept_token_indices(plan), dtype=np.int64)
        assert plan.n_total == len(want) and np.array_equal(got, want), (T, w, p, bud)
        assert np.array_equal(oracle.evict_indices(4, T, window=w, prefix=p, budget=bud).numpy(), want)
        saw_table |= bool(plan.needs_table)
    assert saw_table


def test_invalid_parameters_raise():
    for kw in (dict(strategy=L.STRIDED, stride=0), dict(strategy=L.BLOCK, block=0), dict(strategy=L.BLOCK, block=4, keep=5),
               dict(strategy=L.BLOCK, block=4, keep=0), dict(strategy=L.BUDGET, budget=-1), dict(strategy=L.SLIDING, window=-1)):
        st = kw.pop("strategy")
        with pytest.raises(RuntimeError):
            evict_plan(st, 100, **{"window": 8, **kw})
    import torch
    from kvq_b200 import trim_kv_block_old, trim_kv_budget_old, trim_kv_strided
    x = ((torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4)),)
    with pytest.raises(AssertionError):
        trim_kv_strided(x, 2, 0)          # reference implementations.py:164
    with pytest.raises(AssertionError):
        trim_kv_block_old(x, 2, 0, 1)     # :207
    with pytest.raises(AssertionError):
        trim_kv_block_old(x, 2, 4, 5)     # :208
    with pytest.raises(AssertionError):
        trim_kv_budget_old(x, 2, -1)      # :257
