"""Aggregation algebra (hypothesis): weights sum to one, permutation invariance,
K=1 identity, integer-buffer policy."""
from collections import OrderedDict

import pytest
import torch
from hypothesis import given, settings, strategies as st

from baton_b200.parallel.aggregate import client_weights, fedavg_into, fedavg_loss_history


@given(st.lists(st.integers(min_value=0, max_value=10_000), min_size=1, max_size=16))
def test_weights_sum_to_one_or_zero(ns):
    w = client_weights(ns)
    assert len(w) == len(ns)
    if sum(ns) == 0:
        assert w == [0.0] * len(ns)
    else:
        assert sum(w) == pytest.approx(1.0)
        assert all(x >= 0 for x in w)


def _states(k, seed):
    g = torch.Generator().manual_seed(seed)
    return [OrderedDict(w=torch.randn(5, 3, generator=g), b=torch.randn(3, generator=g),
                        steps=torch.tensor(int(torch.randint(0, 100, (1,), generator=g))))
            for _ in range(k)]


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 8), st.integers(0, 1000), st.data())
def test_fedavg_matches_reference_formula_and_is_permutation_invariant(k, seed, data):
    ns = data.draw(st.lists(st.integers(1, 500), min_size=k, max_size=k))
    states = _states(k, seed)
    glob = OrderedDict((key, torch.zeros_like(v)) for key, v in states[0].items())
    assert fedavg_into(glob, states, ns)
    N = sum(ns)
    for key in ("w", "b"):
        want = sum(s[key] * n for s, n in zip(states, ns)) / N       # manager.py:124-126
        assert torch.allclose(glob[key], want, atol=1e-5)
    assert int(glob["steps"]) == max(int(s["steps"]) for s in states)
    perm = data.draw(st.permutations(list(range(k))))
    glob2 = OrderedDict((key, torch.zeros_like(v)) for key, v in states[0].items())
    fedavg_into(glob2, [states[i] for i in perm], [ns[i] for i in perm])
    for key in glob:
        assert torch.allclose(glob[key].float(), glob2[key].float(), atol=1e-5)


def test_k1_identity_zero_samples_noop_and_int_mean():
    s = _states(1, 3)
    glob = OrderedDict((k, torch.full_like(v, 7)) for k, v in s[0].items())
    assert fedavg_into(glob, s, [10])
    assert torch.allclose(glob["w"], s[0]["w"], atol=1e-6)
    before = {k: v.clone() for k, v in glob.items()}
    assert fedavg_into(glob, s, [0]) is False
    assert all(torch.equal(glob[k], before[k]) for k in glob)
    a = OrderedDict(n=torch.tensor(10)); b = OrderedDict(n=torch.tensor(20))
    g = OrderedDict(n=torch.tensor(0))
    fedavg_into(g, [a, b], [1, 3], int_policy="mean")
    assert int(g["n"]) == 18                    # round(10*.25 + 20*.75) = round(17.5) -> 18
    with pytest.raises(KeyError):
        fedavg_into(OrderedDict(z=torch.zeros(1)), [a], [1])
    with pytest.raises(ValueError):
        fedavg_into(g, [a, b], [1])


def test_inplace_write_into_live_parameters():
    m = torch.nn.Linear(3, 2)
    ptr = m.weight.data_ptr()
    s = [OrderedDict((k, torch.ones_like(v)) for k, v in m.state_dict().items()),
         OrderedDict((k, 3 * torch.ones_like(v)) for k, v in m.state_dict().items())]
    fedavg_into(m.state_dict(), s, [1, 1])
    assert m.weight.data_ptr() == ptr and torch.allclose(m.weight, torch.full_like(m.weight, 2.0))


def test_loss_history_weighting():
    out = fedavg_loss_history([[4.0, 2.0], [1.0, 1.0]], [1, 3], 2)
    assert out == pytest.approx([1.75, 1.25])
    assert fedavg_loss_history([[4.0], [1.0, 1.0]], [1, 3], 2) == pytest.approx([1.75, 1.0])
    assert fedavg_loss_history([], [], 3) == []
