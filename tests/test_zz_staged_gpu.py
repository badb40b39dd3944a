"""GPU checks of the pieces added after this round's last GPU session (written and CPU-tested without a device; the file name
sorts last so that the suite validated on the GPU runs first): the values constant folding produces for the three graphs of
tenncor/hone/test/test_cstrules.cpp, and the DQN replay environment (tenncor_b200/extenncor, reference extenncor/dqn_trainer.py)
end to end — greedy action against a numpy forward pass, training steps, checkpoint and resume."""
import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200 import extenncor
from tests.test_backprop_golden import render_typed, same_graph
from tests.test_hone import CST_A, CST_B, _cstrules_graphs


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


@pytest.mark.gpu
def test_cstrules_goldens(gpu):  # tenncor/hone/test/test_cstrules.cpp: Typical :12-48, StopAtVar :51-90, IdentityDependency :93-130
    lhs, typical, stop_at_var, identity = _cstrules_graphs()
    (got,), _ = tc.optimize([typical])
    assert same_graph(render_typed(got), "(constant:[107\\44\\180\\90\\193\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"), render_typed(got)
    np.testing.assert_array_equal(got.data(), CST_B + 4 + CST_A + CST_B)
    (got,), _ = tc.optimize([stop_at_var])
    assert same_graph(render_typed(got),
                      "(ADD<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(constant:[26\\19\\78\\42\\65\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(ADD<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_____`--(variable:a<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_____`--(constant:[22\\15\\74\\38\\61\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"), render_typed(got)
    (got,), _ = tc.optimize([identity])
    assert same_graph(render_typed(got),
                      "(IDENTITY<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(constant:[26\\19\\78\\42\\65\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"
                      "_`--(constant:[81\\25\\102\\48\\128\\...]<DOUBLE>[2\\3\\4\\1\\1\\1\\1\\1])\n"), render_typed(got)


def _sigmoid(x):
    return 1 / (1 + np.exp(-x))


def _bgd(error, leaves):
    return tc.api.approx.rms_momentum(error, leaves, learning_rate=0.1, discount_factor=0.5, apply=lambda x: tc.api.clip_by_l2norm(x, 5))


def _env(tmp_path, seed=4, **kw):
    tc.seed(seed)
    model = tc.api.layer.link([
        tc.api.layer.dense([10], [9]), tc.api.layer.bind(tc.api.sigmoid),
        tc.api.layer.dense([9], [9]), tc.api.layer.bind(tc.api.sigmoid)])
    args = dict(mbatch_size=8, store_interval=1, train_interval=1, discount_rate=0.99, explore_period=0, action_prob=0.0,
                usecase="g", cachedir=str(tmp_path))
    args.update(kw)
    return extenncor.DQNEnv(model, _bgd, **args), model


@pytest.mark.gpu
def test_dqn_env_end_to_end(gpu, tmp_path):
    env, model = _env(tmp_path)
    w0, b0, w1, b1 = [v.data().astype(np.float64) for v in model.get_storage()]
    rng = np.random.default_rng(0)

    def greedy(obs):
        return int(np.argmax(_sigmoid(_sigmoid(obs @ w0 + b0) @ w1 + b1)))

    for _ in range(5):                                        # exploration probability 0: every action is the network's argmax
        obs = rng.random(10).astype(np.float32)
        assert env.action(obs) == greedy(obs.astype(np.float64))
    for i in range(32):
        obs, nxt = rng.random(10).astype(np.float32), rng.random(10).astype(np.float32)
        env.store(obs, int(rng.integers(0, 9)), float(rng.uniform(-1, 1)), nxt)
    before = [v.data().copy() for v in model.get_storage()]
    errs = [env.train() for _ in range(10)]
    assert all(e is not None and np.isfinite(e) and e >= 0 for e in errs) and env.ntrain_called == 10
    assert any(not np.array_equal(v.data(), b) for v, b in zip(model.get_storage(), before))   # the source net trained

    probe = rng.random(10).astype(np.float32)
    want_action = env.action(probe)
    assert env.backup()
    resumed, _ = _env(tmp_path, seed=99)                       # different fresh weights: the checkpoint must win
    assert resumed.recovered and resumed.ntrain_called == 10 and len(resumed.experiences) == 32
    assert resumed.action(probe) == want_action
    err = resumed.train()
    assert err is not None and np.isfinite(err)
