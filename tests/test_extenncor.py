"""The DQN replay environment and its checkpoint / resume (extenncor/dqn_trainer.py, trainer_cache.py), host side:
graph construction, exploration schedule, replay buffer, mini-batch assembly, the reference's `dqn.DqnEnv` checkpoint message and
a full backup -> new process-like recovery of graphs + weights + buffer. Evaluation itself is a GPU matter
(tests/test_zz_staged_gpu.py)."""
import os

import numpy as np
import pytest

import tenncor_b200 as tc
from tenncor_b200 import configs, extenncor
from tenncor_b200.extenncor import dqn_trainer, trainer_cache
from tests.test_backprop_golden import render_typed


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def make_model(nobs=10, nunits=9, nactions=9):
    return tc.api.layer.link([
        tc.api.layer.dense([nobs], [nunits]), tc.api.layer.bind(tc.api.sigmoid),
        tc.api.layer.dense([nunits], [nactions]), tc.api.layer.bind(tc.api.sigmoid)])


def bgd(error, leaves):  # demo/dqn_demo.py:91-93
    return tc.api.approx.rms_momentum(error, leaves, learning_rate=0.1, discount_factor=0.5, apply=lambda x: tc.api.clip_by_l2norm(x, 5))


def make_env(tmp_path, seed=4, **kw):
    tc.seed(seed)
    args = dict(mbatch_size=4, store_interval=1, train_interval=1, discount_rate=0.99, usecase="t", cachedir=str(tmp_path))
    args.update(kw)
    return extenncor.DQNEnv(make_model(), bgd, **args)


def test_training_graph_is_config_c5(tmp_path):
    """the environment's training step is the graph bench.py times as C5 (configs.dqn), behind one extra IDENTITY"""
    env = make_env(tmp_path, mbatch_size=32)
    cfg = configs.dqn(nbatch=32, discount_rate=0.99)
    got, want = render_typed(env.prediction_err).split("\n"), render_typed(cfg.train).split("\n")
    assert got[0].startswith("(IDENTITY<FLOAT>[1\\1")
    strip = lambda lines: [l.strip("_|` -") for l in lines]  # noqa: E731
    assert strip(got[1:]) == strip(want)
    assert env.act_idx.opname() == "ARGMAX" and env.src_shape == [32, 9] and env.mbatch_size == 32
    assert not env.recovered


def test_linear_annealing_and_exploration(tmp_path):
    env = make_env(tmp_path, explore_period=1000, action_prob=0.05)
    assert env._linear_annealing(1.) == 1.
    env.actions_executed = 500
    assert abs(env._linear_annealing(1.) - 0.525) < 1e-12
    env.actions_executed = 1000
    assert env._linear_annealing(1.) == 0.05
    env.actions_executed = 0
    acts = [env.action(np.zeros(10)) for _ in range(50)]   # exploration probability ~1: random actions, nothing is evaluated
    assert env.actions_executed == 50 and all(0 <= a < 9 for a in acts) and len(set(acts)) > 3


def test_replay_buffer(tmp_path):
    env = make_env(tmp_path, store_interval=2, max_exp=3)
    for i in range(10):
        env.store([float(i)] * 10, i % 9, 0.5 * i, [float(i + 1)] * 10)
    assert env.nstore_called == 10
    assert [e[1] for e in env.experiences] == [4, 6, 8]      # every 2nd call is kept, the oldest beyond max_exp are dropped
    assert env.train() is None and env.ntrain_called == 0     # fewer experiences than a mini-batch: nothing happens


def test_batch_assembly(tmp_path):
    env = make_env(tmp_path)
    samples = [([0.1 * i] * 10, i, 1.0 - i, [0.2 * i] * 10) for i in range(4)]
    states, mask, new_states, rewards = env.assemble_batch(samples)
    assert states.shape == (4, 10) == tuple(env.src_obs.shape()) and new_states.shape == (4, 10)
    np.testing.assert_array_equal(mask, np.eye(9, dtype=np.float32)[:4])
    np.testing.assert_allclose(states[:, 0], [0, 0.1, 0.2, 0.3], rtol=1e-6)
    np.testing.assert_array_equal(rewards, [1, 0, -1, -2])
    assert mask.shape == tuple(env.src_outmask.shape()) and rewards.shape == tuple(env.rewards.shape())


def test_env_message_is_the_reference_protobuf():
    """encode_env / decode_env against google.protobuf on a dynamically built copy of extenncor/dqn_trainer.proto"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="dqn_trainer_mirror.proto", package="dqn", syntax="proto3")
    exp = fd.message_type.add(name="ExpBatch")
    F = descriptor_pb2.FieldDescriptorProto
    exp.field.add(name="act_idx", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    exp.field.add(name="reward", number=2, type=F.TYPE_FLOAT, label=F.LABEL_OPTIONAL)
    exp.field.add(name="obs", number=3, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED)
    exp.field.add(name="new_obs", number=4, type=F.TYPE_FLOAT, label=F.LABEL_REPEATED)
    env = fd.message_type.add(name="DqnEnv")
    env.field.add(name="actions_executed", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    env.field.add(name="ntrain_called", number=2, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    env.field.add(name="nstore_called", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    env.field.add(name="experiences", number=4, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name=".dqn.ExpBatch")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    DqnEnv = message_factory.GetMessageClass(pool.FindMessageTypeByName("dqn.DqnEnv"))

    experiences = [([0.5, 1.5, -2.0], 3, 0.25, [1.0, 2.0, 3.0]), ([0.0] * 3, 0, 0.0, [4.0, 5.0, 6.0]), ([7.0], 300, -1.5, [8.0])]
    data = dqn_trainer.encode_env(1234, 0, 77, experiences)
    msg = DqnEnv()
    msg.ParseFromString(data)                                 # what we write, the reference's message class reads
    assert (msg.actions_executed, msg.ntrain_called, msg.nstore_called) == (1234, 0, 77)
    assert [(list(e.obs), e.act_idx, e.reward, list(e.new_obs)) for e in msg.experiences] == experiences
    assert msg.SerializeToString() == data                    # byte for byte the canonical proto3 encoding
    msg.actions_executed = -5                                 # and what protobuf writes, we read (negative int32 = 10-byte varint)
    assert dqn_trainer.decode_env(msg.SerializeToString()) == (-5, 0, 77, experiences)


def test_backup_and_recovery(tmp_path):
    env = make_env(tmp_path)
    for i in range(6):
        env.store(list(np.full(10, 0.1 * i, dtype=np.float32)), i, 0.5 * i, list(np.full(10, 0.2 * i, dtype=np.float32)))
    env.actions_executed, env.ntrain_called = 42, 7
    weights = [v.data().copy() for v in _trainables(env.prediction_err)]
    assert len(weights) == 12                                 # 2 nets x (2 weights + 2 biases) + 4 rms momentum slots
    tc.to_variable(env.src_obs).assign(np.full((4, 10), 0.75, dtype=np.float32))
    assert env.backup() and env.env_id == 1 and env.session_cache.cur_id == 1
    assert sorted(os.listdir(os.path.join(str(tmp_path), "t", "dqn"))) == ["env_1.bkup", "session_1.onnx"]

    other = make_env(tmp_path, seed=99)  # a different process: different fresh weights — recovery must bring the saved ones back
    assert other.recovered and other.env_id == 1
    assert (other.actions_executed, other.ntrain_called, other.nstore_called) == (42, 7, 6)
    assert [(e[1], e[2]) for e in other.experiences] == [(e[1], e[2]) for e in env.experiences]
    np.testing.assert_allclose(other.experiences[3][0], env.experiences[3][0])
    assert render_typed(other.prediction_err) == render_typed(env.prediction_err)
    assert render_typed(other.act_idx) == render_typed(env.act_idx)
    for got, want in zip([v.data() for v in _trainables(other.prediction_err)], weights):
        np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(other.src_obs.data(), np.full((4, 10), 0.75, dtype=np.float32))
    assert other.src_shape == [4, 9] and other.mbatch_size == 4
    # the recovered handles are live parts of the recovered graphs, not copies
    assert any(leaf == other.src_obs for leaf in _leaves(other.prediction_err))
    assert other.backup() and other.env_id == 2 and other.session_cache.cur_id == 2   # numbering continues

    fresh = make_env(tmp_path, seed=99, clean_startup=True)   # clean: ignore what is there
    assert any(not np.array_equal(v.data(), w) for v, w in zip(_trainables(fresh.prediction_err), weights))
    assert not fresh.recovered and fresh.experiences == [] and fresh.actions_executed == 0
    assert trainer_cache._id_cachefile(os.path.join(str(tmp_path), "t", "dqn", "session_2.onnx"), "session_", ".onnx") == 2
    assert trainer_cache._id_cachefile(os.path.join(str(tmp_path), "t", "dqn", "env_1.bkup"), "session_", ".onnx") is None


def _trainables(root):
    return [t for t in _leaves(root) if t.usage() == "variable" and str(t) in ("weight", "bias", "momentum")]


def _leaves(root):
    seen, out, stack = set(), [], [root]
    while stack:
        t = stack.pop()
        if hash(t) in seen:
            continue
        seen.add(hash(t))
        if t.is_leaf():
            out.append(t)
        stack.extend(t.args())
    return out
