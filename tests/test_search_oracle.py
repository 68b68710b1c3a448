"""CPU: known answers of the whole search driver.  The numbers below were produced by the REAL reference
(Cython-compiled ctree + the reference's Python SampledMCTS + its unit-test MockModel)
and pin oracle/search_oracle.py + the CPU tree end to end."""
import numpy as np
import pytest

from _mock import MockConfig, MockModel, sequential_search

KAT = [  # (root value, marginal visit counts, root children) per agent turn; final joint action [[0, 1]]
    (0.5862223, [[[1, 1, 1]]], [[0], [1], [2]]),
    (0.63890105, [[[0, 2, 1]]], [[1], [2]]),
]


def kat_legal():
    legal = np.ones((1, 2, 3), dtype=np.int32)
    legal[0, 1, 0] = 0  # unit_test_mcts.py:171-172
    return legal


def check_kat(outs, chosen):
    for o, (val, mv, children) in zip(outs, KAT):
        assert o.value.shape == (1,) and o.marginal_visit_count.shape == (1, 1, 3)  # the reference's own asserts
        assert o.sampled_actions[0].shape[1] == 1
        assert abs(float(o.value[0]) - val) < 1e-7
        assert o.marginal_visit_count.tolist() == mv
        assert o.sampled_actions[0].tolist() == children
    assert chosen.tolist() == [[0, 1]]


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_reference_unit_test_known_answers(oracle_built, kind):
    from oracle.search_oracle import reference_batch_search

    if not oracle_built.available(kind):
        pytest.skip(f"{kind} tree not built")
    cfg, model, rs = MockConfig(), MockModel(2, 3), np.random.RandomState(123)
    fn = lambda m, o, k, f, n, l: reference_batch_search(cfg, rs, m, o, k, f, n, l, "cpu", add_noise=True, tree_kind=kind)
    check_kat(*sequential_search(fn, cfg, model, legal=kat_legal()))
