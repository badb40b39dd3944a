"""hone pre-evaluation rewrites (tenncor/hone/src/duplicates.cpp, cstrules.hpp, optimize.cpp):
duplicate merging is pure host logic and is checked here on CPU — the rewritten graph must evaluate (CPU oracle) to
exactly the values of the original; constant folding evaluates on the device and is checked under -m gpu (the values of the three test_cstrules.cpp graphs in
tests/test_zz_staged_gpu.py)."""
import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc
from tenncor_b200 import configs


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def _eval(roots):
    tape = tc.dump_graph(roots)
    ids = tc.dump_ids(roots, tape)
    for node in tape:
        if node["kind"] == "leaf":
            node["data"] = np.array(node["data"], copy=True)
    vals = orc.eval_tape(tape)
    return [np.asarray(vals[ids[r]]).reshape(-1).copy() for r in roots], len(tape)


def test_structurally_equal_subgraphs_get_one_owner():
    rng = np.random.default_rng(0)
    a = tc.variable(rng.random((3, 4), dtype=np.float32), "a")
    b = tc.variable(rng.random((3, 4), dtype=np.float32), "b")
    left = tc.api.sigmoid(a * b + 2.0)
    right = tc.api.sigmoid(2.0 + b * a)  # commutative operands in the other order (duplicates.hpp:64-68)
    other = tc.api.sigmoid(a * b + 3.0)
    root = left * right + other
    want, n_before = _eval([root])
    (merged,), removed = tc.merge_dups([root])
    got, n_after = _eval([merged])
    np.testing.assert_array_equal(got[0], want[0])
    assert removed > 0 and n_after < n_before
    # left and right collapsed: the product's two arguments are now the same node
    prod = merged.args()[0]
    assert prod.opname() == "MUL" and prod.args()[0] is prod.args()[1] or prod.args()[0] == prod.args()[1]
    # `other` differs by a constant and must survive
    assert merged.args()[1].opname() == "SIGMOID" and merged.args()[1] != prod.args()[0]


def test_attributes_and_shapes_distinguish_nodes():
    rng = np.random.default_rng(1)
    x = tc.variable(rng.random((2, 3, 4), dtype=np.float32), "x")
    r0, r1 = tc.api.reduce_sum(x, 0, 1), tc.api.reduce_sum(x, 1, 1)
    p0, p1 = tc.api.permute(x, [1, 0, 2]), tc.api.permute(x, [1, 0, 2])
    roots = [r0, r1, p0, p1]
    want, _ = _eval(roots)
    merged, removed = tc.merge_dups(roots)
    assert removed == 1 and merged[2] == merged[3] and merged[0] != merged[1]
    got, _ = _eval(merged)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_variables_and_non_idempotent_nodes_are_never_merged():
    v1 = tc.variable(np.ones((2, 2), dtype=np.float32), "v")
    v2 = tc.variable(np.ones((2, 2), dtype=np.float32), "v")  # equal data, distinct storage
    lo, hi = tc.scalar_constant(0, [2, 2]), tc.scalar_constant(1, [2, 2])
    r1, r2 = tc.api.random.rand_unif(lo, hi), tc.api.random.rand_unif(lo, hi)
    roots = [v1 + r1, v2 + r2, tc.api.assign_add(v1, lo), tc.api.assign_add(v1, lo)]
    merged, removed = tc.merge_dups(roots)
    assert merged[0] != merged[1] and merged[2] != merged[3]
    assert merged[0].args()[1] != merged[1].args()[1]  # two independent draws stay two nodes (deviation noted in hone.hpp)


@pytest.mark.parametrize("build", [lambda: configs.mlp(10, 9, 5, 3), lambda: configs.dqn(nbatch=8), lambda: configs.cnn(),
                                   lambda: configs.recurrent("gru", vocab=6, hidden=5, seq=4, batch=3)], ids=["mlp", "dqn", "cnn", "gru"])
def test_training_graphs_keep_their_values(build):
    cfg = build()
    rng = np.random.default_rng(2)
    for f in cfg.feeds.values():
        f.assign(rng.random(f.shape(), dtype=np.float32) + 0.1)
    want, n_before = _eval([cfg.train])
    # the oracle's ASSIGNs mutated only its own copy of the leaves: the graph still holds the initial weights
    (root,), removed = tc.merge_dups([cfg.train])
    got, n_after = _eval([root])
    np.testing.assert_array_equal(got[0], want[0])
    assert removed > 0 and n_after < n_before
    assert len(tc.describe_plan([root])) <= len(tc.describe_plan([build().train]))


CST_A = np.array([59, 10, 28, 10, 67, 62, 23, 4, 55, 77, 28, 16, 82, 52, 47, 16, 7, 85, 37, 2, 8, 52, 62, 43], dtype=np.float64).reshape(4, 3, 2)
CST_B = np.array([22, 15, 74, 38, 61, 95, 62, 81, 99, 76, 7, 22, 56, 50, 19, 13, 12, 10, 31, 40, 60, 54, 6, 83], dtype=np.float64).reshape(4, 3, 2)


def _cstrules_graphs():
    """the three graphs of tenncor/hone/test/test_cstrules.cpp (Typical :12-48, StopAtVar :51-90, IdentityDependency :93-130)"""
    a, b = tc.constant(CST_A), tc.constant(CST_B)
    c = tc.scalar_constant(4, [4, 3, 2], "DOUBLE")
    var = tc.variable(CST_A, "a")
    lhs = b + c
    typical = lhs + (a + b)
    stop_at_var = lhs + (var + b)
    identity = tc.egen.make_functor("IDENTITY", [lhs, a + b])
    return lhs, typical, stop_at_var, identity


def test_which_functors_fold():
    """the host-side half of constant folding, no device: WHICH functors get evaluated and replaced"""
    lhs, typical, stop_at_var, identity = _cstrules_graphs()
    assert tc.fold_candidates([typical]) == [typical]              # constants all the way down: the root itself becomes the constant
    assert tc.fold_candidates([stop_at_var]) == [lhs]              # a variable below stops the chain: only the constant side folds
    # IDENTITY is never a folding candidate (tenncor/hone/src/cstrules.cpp:26): its arguments fold, it stays, and so does its reader
    assert tc.fold_candidates([identity]) == identity.args()
    assert tc.fold_candidates([tc.api.sin(identity)]) == identity.args()
    lo, hi = tc.scalar_constant(0, [2, 2]), tc.scalar_constant(1, [2, 2])
    noise = tc.api.random.rand_unif(lo, hi)
    # deviation kept on purpose: a random draw is not a constant (only the broadcast of the scalar 2 folds here)
    assert [t.opname() for t in tc.fold_candidates([noise * 2.0])] == ["EXTEND"]


@pytest.mark.gpu
def test_constant_folding_on_device(gpu):
    rng = np.random.default_rng(3)
    x = tc.variable(rng.random((4, 5), dtype=np.float32), "x")
    c1 = tc.constant(np.arange(20, dtype=np.float32).reshape(4, 5))
    c2 = tc.constant(np.full((4, 5), 0.5, dtype=np.float32))
    folded_part = tc.api.exp(c2) * c1 + tc.api.extend_like(tc.api.reduce_sum(c1), c1)  # constants only: becomes one constant leaf
    lo, hi = tc.scalar_constant(0, [4, 5]), tc.scalar_constant(1, [4, 5])
    noise = tc.api.random.rand_unif(lo, hi)                            # constant arguments, but not foldable
    root = x * folded_part + noise * 0.0
    want = x.data() * (np.exp(np.float32(0.5)) * np.arange(20, dtype=np.float32).reshape(4, 5) + np.float32(190))
    (opt,), stats = tc.optimize([root])
    assert stats["folded"] >= 1 and stats["functors_after"] < stats["functors_before"], stats
    np.testing.assert_allclose(opt.get(), want, rtol=1e-5)
    mul = opt.args()[0]
    assert mul.opname() == "MUL" and mul.args()[1].opname() == ""  # the constant sub-graph is a leaf now
    ops = []

    def walk(t):
        ops.append(t.opname())
        for a in t.args():
            walk(a)
    walk(opt)
    assert "RAND_UNIF" in ops and "EXP" not in ops and "REDUCE_SUM" not in ops


@pytest.mark.gpu
def test_optimized_training_step_matches_oracle(gpu):
    from tests.test_train_gpu import OracleSession, rel_err
    cfg = configs.dqn(nbatch=32)
    (train,), stats = tc.optimize([cfg.train])
    assert stats["merged"] > 0
    sess = OracleSession([train])
    rng = np.random.default_rng(4)
    for step in range(3):
        batch = configs.dqn_batch(rng, cfg.feeds)
        for k, f in cfg.feeds.items():
            f.assign(batch[k])
            sess.assign(f, batch[k])
        assert rel_err(train.get(), sess.run()[0]) < 1e-4
