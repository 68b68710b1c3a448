"""CPU, build container only (needs /root/reference): pins oracle/search_oracle.py against the reference's OWN Python
driver.  The unmodified `core/mcts/tree_search/mcts_sampled.py` is imported (ray / gymnasium stubbed, as in
tests/golden/make_golden_model.py) with the compiled reference tree injected as its `cytree` module, and must give
exactly the results of the restated driver for the same model, inputs and RandomState -- in every sequential-agent
turn, with a legal-action mask and exploration noise."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("MAZ_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "core", "mcts", "tree_search", "mcts_sampled.py")),
                                reason="reference sources not present on this box")


@pytest.fixture(scope="module")
def reference_mcts(oracle_built):
    if not oracle_built.available("reference"):
        pytest.skip("oracle/_ref/libmazref.so not built")
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_model as mg

    mg.install_stubs()
    # the reference imports `from core.mcts.ctree.ctree_sampled import cytree` (a Cython module): inject the compiled
    # reference C++ behind the same class name instead of building the Cython extension
    shim = types.ModuleType("core.mcts.ctree.ctree_sampled.cytree")
    shim.Tree_batch = lambda *a: oracle_built.OracleTreeBatch(*a, kind="reference")
    pkg = types.ModuleType("core.mcts.ctree.ctree_sampled")
    pkg.__path__ = []
    pkg.cytree = shim
    for name in ("core.mcts.ctree", "core.mcts.ctree.ctree_sampled"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["core.mcts.ctree.ctree_sampled"] = pkg
    sys.modules["core.mcts.ctree.ctree_sampled.cytree"] = shim
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_mcts_sampled", os.path.join(REF, "core", "mcts", "tree_search", "mcts_sampled.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mg


def test_restated_driver_equals_reference_driver(reference_mcts):
    from oracle.search_oracle import reference_batch_search

    mod, mg = reference_mcts
    from _mock import MockConfig

    N, A, B, K, S = 3, 9, 12, 5, 12
    model = mg.reference_model(N, A, 32)          # the REAL reference MAMuZeroNet (small hidden size), eval mode
    cfg = MockConfig(N, A, S, K)
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(B, N, 16, 1, 1, generator=g)
    with torch.no_grad():
        out0 = model.initial_inference(obs)
    legal = (np.random.RandomState(3).rand(B, N, A) < 0.7).astype(np.float32)
    legal[..., 1] = 1
    factor = np.random.RandomState(4).randint(0, A, size=(B, N)).astype(np.int32)
    for cur in range(N):
        ref = mod.SampledMCTS(cfg, np.random.RandomState(7)).batch_search(
            model, out0, cur, factor, N, legal.copy(), torch.device("cpu"), add_noise=True, sampled_tau=1.0)
        mine = reference_batch_search(cfg, np.random.RandomState(7), model, out0, cur, factor, N, legal.copy(), torch.device("cpu"),
                                      add_noise=True, sampled_tau=1.0, tree_kind="reference")
        assert np.array_equal(ref.value, mine.value)
        assert np.array_equal(ref.marginal_visit_count, mine.marginal_visit_count)
        assert np.array_equal(ref.marginal_priors, mine.marginal_priors)
        for f in ("sampled_actions", "sampled_visit_count", "sampled_qvalues", "sampled_imp_ratio", "sampled_priors"):
            for b in range(B):
                assert np.array_equal(getattr(ref, f)[b], getattr(mine, f)[b]), (cur, f, b)
