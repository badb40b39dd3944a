"""Replay-buffer DQN environment over the evaluation path (reference: extenncor/dqn_trainer.py:1-243; config C5).

`DQNEnv` owns two graphs: `act_idx` = ARGMAX of the source network on one observation (environment interaction) and
`prediction_err` = one training step (masked TD error of the source network against the target network's best next value,
the caller's update rule on the source weights, soft update of the target weights — all in one apply_update root, evaluated as
one planned launch sequence). `action / store / train` are the reference's loop; `backup()` writes a checkpoint (graphs with
their weights and optimiser state + counters and replay buffer) and a new `DQNEnv` on the same `usecase` resumes from it.
The `.bkup` file is the reference's `dqn.DqnEnv` protobuf message (extenncor/dqn_trainer.proto), written and read here by a
small wire-format codec, so checkpoints of either side's replay buffer are interchangeable."""
import math
import os
import random
import struct

import numpy as np

import tenncor_b200 as tc

from . import trainer_cache as ecache

_get_random = None


def _uniform():
    global _get_random
    if _get_random is None:
        _get_random = tc.unif_gen(0, 1)  # the seeded host generator, like the reference's module-level tc.unif_gen(0, 1)
    return _get_random()


def get_dqnupdate(update_fn, update_rate):
    """source weights by `update_fn`; target weights move `update_rate` of the way to the UPDATED source weights (:15-29)"""
    def dqnupdate(err, variables):
        half = len(variables) // 2
        src_vars, nxt_vars = variables[:half], variables[half:]
        src_updates = update_fn(err, src_vars)
        updated = {var: upd for var, upd in src_updates}
        assigns = []
        for nxt_var, src_var in zip(nxt_vars, src_vars):
            diff = nxt_var - updated[src_var]
            assigns.append((nxt_var, tc.api.assign_sub(nxt_var, update_rate * diff)))
        return assigns
    return dqnupdate


def get_dqnerror(env, discount_rate):
    """mean squared TD error of the taken action's score (:31-48)"""
    def dqnerror(models):
        src_model, nxt_model = tuple(models)
        src_act = src_model.connect(env.src_obs)                       # forward action score computation
        nxt_act = nxt_model.connect(env.nxt_obs)                       # predicting target future rewards
        target_vals = env.nxt_outmask * tc.api.reduce_max_1d(nxt_act, 0)
        future_reward = env.rewards + discount_rate * target_vals
        masked_output_score = tc.api.reduce_sum_1d(src_act * env.src_outmask, 0)
        return tc.api.reduce_mean(tc.api.square(masked_output_score - future_reward))
    return dqnerror


# ---------------------------------------------------------------- dqn.DqnEnv on the wire (extenncor/dqn_trainer.proto, proto3)
def _varint(n):
    n &= (1 << 64) - 1  # negative int32 are sign-extended to 64 bits
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return val, pos


def _int32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _fields(buf):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = _read_varint(buf, pos)
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            val, pos = bytes(buf[pos:pos + n]), pos + n
        elif wire == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        elif wire == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        else:
            raise ValueError("unsupported wire type %d" % wire)
        yield field, wire, val


def _floats(wire, val):  # repeated float: packed (proto3 default) or one fixed32 per element
    return list(struct.unpack("<%df" % (len(val) // 4), val)) if wire in (2, 5) else []


def encode_env(actions_executed, ntrain_called, nstore_called, experiences):
    out = bytearray()
    for field, v in ((1, actions_executed), (2, ntrain_called), (3, nstore_called)):
        if v:  # proto3 omits defaults
            out += _varint(field << 3) + _varint(int(v))
    for obs, act_idx, reward, new_obs in experiences:
        exp = bytearray()
        if act_idx:
            exp += _varint(1 << 3) + _varint(int(act_idx))
        if reward:
            exp += _varint((2 << 3) | 5) + struct.pack("<f", float(reward))
        for field, vec in ((3, obs), (4, new_obs)):
            vec = np.asarray(vec, dtype=np.float32).reshape(-1)
            if vec.size:
                exp += _varint((field << 3) | 2) + _varint(4 * vec.size) + vec.astype("<f4").tobytes()
        out += _varint((4 << 3) | 2) + _varint(len(exp)) + bytes(exp)
    return bytes(out)


def decode_env(data):
    counters = {1: 0, 2: 0, 3: 0}
    experiences = []
    for field, wire, val in _fields(data):
        if field in counters and wire == 0:
            counters[field] = _int32(val)
        elif field == 4 and wire == 2:
            act_idx, reward, obs, new_obs = 0, 0.0, [], []
            for f, w, v in _fields(val):
                if f == 1 and w == 0:
                    act_idx = _int32(v)
                elif f == 2 and w == 5:
                    reward = struct.unpack("<f", v)[0]
                elif f == 3:
                    obs += _floats(w, v)
                elif f == 4:
                    new_obs += _floats(w, v)
            experiences.append((obs, act_idx, reward, new_obs))
    return counters[1], counters[2], counters[3], experiences


# ---------------------------------------------------------------- the environment
class DQNEnv(ecache.EnvManager):
    def __init__(self, src_model, update_fn, optimize_cfg="", max_exp=30000, train_interval=5, store_interval=5, explore_period=1000,
                 action_prob=0.05, mbatch_size=32, discount_rate=0.95, target_update_rate=0.01, clean_startup=False,
                 usecase="", cachedir="/tmp", ctx=None):
        """same parameters as the reference's DQNEnv (:52-60). `optimize_cfg`: the reference runs its json rule file over the context
        (tc.optimize(cfg, ctx)); that file is not part of the repository, so a non-empty value asks for the code-defined hone passes
        (duplicate merging, tc.optimize; host-side only) over the two graphs instead. `ctx` is accepted and unused (one context)."""
        self.max_exp = max_exp
        self.train_interval = train_interval
        self.store_interval = store_interval
        self.explore_period = explore_period
        self.action_prob = action_prob

        def default_init():
            self.actions_executed = 0
            self.ntrain_called = 0
            self.nstore_called = 0
            self.experiences = []
            nxt_model = src_model.deep_clone()
            inshape = list(src_model.get_input().shape())
            batchin = [mbatch_size] + inshape
            # environment interaction
            self.obs = tc.EVariable(inshape, 0, "obs")
            self.act_idx = tc.api.argmax(src_model.connect(self.obs))
            # training
            self.src_obs = tc.EVariable(batchin, 0, "src_obs")
            self.nxt_obs = tc.EVariable(batchin, 0, "nxt_obs")
            self.src_outmask = tc.EVariable([mbatch_size] + list(src_model.shape()), 1, "src_outmask")
            self.nxt_outmask = tc.EVariable([mbatch_size], 1, "nxt_outmask")
            self.rewards = tc.EVariable([mbatch_size], 0, "rewards")
            self.prediction_err = tc.api.identity(tc.apply_update(
                [src_model, nxt_model], get_dqnupdate(update_fn, target_update_rate), get_dqnerror(self, discount_rate)))
            if optimize_cfg:
                (self.act_idx, self.prediction_err), _ = tc.optimize([self.act_idx, self.prediction_err], fold_constants=False)

        super().__init__(os.path.join(usecase, "dqn"), default_init=default_init, clean=clean_startup, cacheroot=cachedir)
        self.src_shape = self.src_outmask.shape()
        mb = self.rewards.shape()
        self.mbatch_size = mb[0] if len(mb) > 0 else 1

    # ---- checkpoint contract (trainer_cache.EnvManager)
    _LEAVES = ("obs", "src_obs", "nxt_obs", "src_outmask", "nxt_outmask", "rewards")

    def roots(self):
        return [self.act_idx, self.prediction_err]

    def handles(self):
        out = {name: getattr(self, name) for name in self._LEAVES}
        out["act_idx"], out["prediction_err"] = self.act_idx, self.prediction_err
        return out

    def restore(self, handles):
        missing = [name for name in self._LEAVES + ("act_idx", "prediction_err") if name not in handles]
        if missing:
            raise RuntimeError("session file lacks %s" % ", ".join(missing))
        for name in self._LEAVES:
            setattr(self, name, tc.to_variable(handles[name]))
        self.act_idx, self.prediction_err = handles["act_idx"], handles["prediction_err"]

    def _backup_env(self, fpath):
        with open(fpath, "wb") as envfile:
            envfile.write(encode_env(self.actions_executed, self.ntrain_called, self.nstore_called, self.experiences))
        return True

    def _recover_env(self, fpath):
        with open(fpath, "rb") as envfile:
            self.actions_executed, self.ntrain_called, self.nstore_called, self.experiences = decode_env(envfile.read())
        return True

    # ---- the loop (dqn_trainer.py:183-243)
    def action(self, obs):
        self.actions_executed += 1
        exploration = self._linear_annealing(1.)
        if _uniform() < exploration:  # perform random exploration action
            return math.floor(_uniform() * self.src_shape[-1])
        self.obs.assign(np.asarray(obs, dtype=np.float32).reshape(self.obs.shape()))
        return int(np.asarray(self.act_idx.get()).reshape(-1)[0])

    def store(self, observation, act_idx, reward, new_obs):
        if 0 == self.nstore_called % self.store_interval:
            self.experiences.append((observation, act_idx, reward, new_obs))
            if len(self.experiences) > self.max_exp:
                self.experiences = self.experiences[1:]
        self.nstore_called += 1

    def assemble_batch(self, samples):
        """(states, action one-hot mask, new states, rewards) of a list of experiences, shaped like the feed variables"""
        nactions = self.src_shape[-1]
        states, new_states, action_mask, rewards = [], [], [], []
        for observation, act_idx, reward, new_obs in samples:
            assert len(new_obs) > 0
            states.append(np.asarray(observation, dtype=np.float32).reshape(-1))
            mask = [0.] * nactions
            mask[act_idx] = 1.
            action_mask.append(mask)
            rewards.append(reward)
            new_states.append(np.asarray(new_obs, dtype=np.float32).reshape(-1))
        return (np.array(states, dtype=np.float32).reshape(self.src_obs.shape()), np.array(action_mask, dtype=np.float32),
                np.array(new_states, dtype=np.float32).reshape(self.nxt_obs.shape()), np.array(rewards, dtype=np.float32))

    def train(self):
        """every `train_interval`-th call: sample a mini-batch from the buffer and run one training step; returns its error"""
        if len(self.experiences) < self.mbatch_size:
            return None
        err = None
        if 0 == (self.ntrain_called % self.train_interval):
            states, action_mask, new_states, rewards = self.assemble_batch(self._random_sample())
            self.src_obs.assign(states)
            self.src_outmask.assign(action_mask)
            self.nxt_obs.assign(new_states)
            self.rewards.assign(rewards)
            err = float(np.asarray(self.prediction_err.get()).reshape(-1)[0])
        self.ntrain_called += 1
        return err

    def _linear_annealing(self, initial_prob):
        if self.actions_executed >= self.explore_period:
            return self.action_prob
        return initial_prob - self.actions_executed * (initial_prob - self.action_prob) / self.explore_period

    def _random_sample(self):
        return random.sample(self.experiences, self.mbatch_size)
