"""The reference's demo models, built through the public API exactly like the demos do.

Each builder returns a `Config` whose `train` node evaluates one full training step
(forward + derivative graph + optimiser ASSIGNs, trainer::apply_update semantics) and whose
`feeds` are the input variables a step assigns. Shapes follow SURVEY.md §8(d) / BASELINE.md §3.
"""
from collections import namedtuple

import numpy as np

import tenncor_b200 as tc

Config = namedtuple("Config", "name train feeds model variables flops_per_step desc")


def mlp(ninput=10, nhidden=9, noutput=5, nbatch=3, learning_rate=0.9, seed=0, name="C1", pixels=False):
    """dense -> sigmoid -> dense -> sigmoid, MSE, SGD (demo/gd_demo.py:55-71, demo/gd_demo.cpp:100-168).
    pixels=True: the input variable holds UINT8 pixels (how MNIST is stored) and the graph starts with the reference's own
    CAST to float and a scale by 1/255 (tenncor/eteq/caster.hpp:10-44): a quarter of the host -> device bytes per step."""
    tc.seed(seed)
    if pixels:
        feed_input = tc.EVariable([nbatch, ninput], 0, "train_input", dtype="uint8")
        train_input = tc.api.cast(feed_input, "float32") * (1.0 / 255.0)
    else:
        feed_input = train_input = tc.EVariable([nbatch, ninput], 0, "train_input")
    train_exout = tc.EVariable([nbatch, noutput], 0, "train_exout")
    model = tc.api.layer.link([
        tc.api.layer.dense([ninput], [nhidden]),
        tc.api.layer.bind(tc.api.sigmoid),
        tc.api.layer.dense([nhidden], [noutput]),
        tc.api.layer.bind(tc.api.sigmoid),
    ], train_input)
    train = tc.apply_update(
        [model],
        lambda err, leaves: tc.api.approx.sgd(err, leaves, learning_rate=learning_rate),
        lambda models: tc.api.loss.mean_squared(train_exout, models[0].connect(train_input)))
    # grad wrt the first layer's input is not built (internal/teq/src/derive.cpp:151-159)
    flops = 4 * nbatch * ninput * nhidden + 6 * nbatch * nhidden * noutput
    return Config(name, train, {"x": feed_input, "y": train_exout}, model, model.get_storage(), flops,
                  "MLP %d-%d-%d sigmoid, MSE, SGD %.2g, batch %d%s" % (ninput, nhidden, noutput, learning_rate, nbatch,
                                                                      ", uint8 pixel input cast on the device" if pixels else ""))


def mlp_batch(rng, cfg_feeds, one_hot=False):
    """synthetic batch: x ~ U[0,1); y = pairwise mean of x (gd_demo batch_generate) or one-hot."""
    x_shape = cfg_feeds["x"].shape()
    y_shape = cfg_feeds["y"].shape()
    if cfg_feeds["x"].dtype() == np.uint8:
        x = rng.integers(0, 256, x_shape, dtype=np.uint8)
    else:
        x = rng.random(x_shape, dtype=np.float32)
    if one_hot or x_shape[1] != 2 * y_shape[1]:
        y = np.zeros(y_shape, dtype=np.float32)
        y[np.arange(y_shape[0]), rng.integers(0, y_shape[1], y_shape[0])] = 1
    else:
        y = ((x[:, 0::2] + x[:, 1::2]) / 2).astype(np.float32)
    return x, y


def encoded_loss(encoded_expect, encoded_result):
    # demo/lstm/latin_demo.py:39-40
    return tc.api.reduce_sum(-tc.api.log(tc.api.reduce_sum(encoded_result * encoded_expect, 0, 1)))


def recurrent(kind="lstm", vocab=128, hidden=1024, seq=128, batch=None, learning_rate=0.1, seed=3, name="C4"):
    """layer.lstm / layer.gru -> dense -> softmax(dim 0), summed NLL, adagrad
    (demo/lstm/latin_demo.py:94-122; GRU twin demo/gru/latin_demo.py). batch=None is the demo's
    literal un-batched form ([N, seq]); a batch puts B at rank 1 and the sequence at rank 2 so
    that every per-step SLICE stays a zero-copy view."""
    tc.seed(seed)
    winit = tc.api.init.random_uniform(-0.05, 0.05)
    make = tc.api.layer.lstm if kind == "lstm" else tc.api.layer.gru
    if batch is None:
        cell = make([vocab], hidden, seq, kernel_init=winit)
        in_shape = [seq, vocab]
    else:
        cell = make([batch, vocab], hidden, seq, kernel_init=winit, seq_dim=2)
        in_shape = [seq, batch, vocab]
    model = tc.api.layer.link([
        cell,
        tc.api.layer.dense([hidden], [vocab], kernel_init=winit),
        tc.api.layer.bind(lambda x: tc.api.softmax(x, 0, 1)),
    ])
    train_inps = tc.EVariable(in_shape, 0, "train_inps")
    train_exout = tc.EVariable(in_shape, 0, "train_exout")
    train = tc.apply_update(
        [model],
        lambda error, leaves: tc.api.approx.adagrad(error, leaves, learning_rate=learning_rate, epsilon=1e-8),
        lambda models: encoded_loss(train_exout, models[0].connect(train_inps)))
    B = batch or 1
    ngates = 4 if kind == "lstm" else 3
    flops = 3 * (seq * ngates * 2 * B * (vocab + hidden) * hidden + 2 * seq * B * hidden * vocab)
    return Config(name, train, {"x": train_inps, "y": train_exout}, model, model.get_storage(), flops,
                  "%s N=%d H=%d seq=%d batch=%s, softmax NLL, adagrad %.2g" % (kind, vocab, hidden, seq, batch, learning_rate))


def recurrent_batch(rng, cfg_feeds, vocab):
    """one-hot token ids ~ U{0..vocab-1}; the target is the input shifted by one step."""
    shape = cfg_feeds["x"].shape()  # [seq, vocab] or [seq, batch, vocab]
    lead = shape[:-1]
    ids = rng.integers(0, vocab, [lead[0] + 1] + list(lead[1:]))
    eye = np.eye(vocab, dtype=np.float32)
    return eye[ids[:-1]], eye[ids[1:]]


def rbm(nvisible=784, nhidden=64, nbatch=4096, learning_rate=0.01, discount=0.95, seed=1, name="C2"):
    """Bernoulli RBM trained by CD-1 (demo/rbm_demo.py:63-88, tenncor/trainer/rbm.hpp:85-165)."""
    tc.seed(seed)
    model = tc.api.layer.rbm(nvisible, nhidden)
    visible = tc.EVariable([nbatch, nvisible], 0, "visible")
    train = tc.rbm_train(model, visible, learning_rate=learning_rate, discount_factor=discount)
    flops = 5 * 2 * nbatch * nvisible * nhidden
    variables = list({id(v): v for v in model.fwd().get_storage() + model.bwd().get_storage()}.values())
    return Config(name, train, {"x": visible}, model, variables, flops,
                  "RBM %d<->%d CD-1, batch %d" % (nvisible, nhidden, nbatch))


def dqn(nobs=10, nunits=9, nactions=9, nbatch=4096, discount_rate=0.99, target_update_rate=0.01,
        learning_rate=0.1, rms_discount=0.5, clip=5.0, seed=4, name="C5"):
    """DQN training step: source + target copies of dense-sigmoid-dense-sigmoid, masked TD error,
    rms_momentum with l2-norm clipping on the source net, soft update of the target net
    (demo/dqn_demo.py:70-93, extenncor/dqn_trainer.py:15-48,72-92)."""
    tc.seed(seed)
    src_model = tc.api.layer.link([
        tc.api.layer.dense([nobs], [nunits]),
        tc.api.layer.bind(tc.api.sigmoid),
        tc.api.layer.dense([nunits], [nactions]),
        tc.api.layer.bind(tc.api.sigmoid),
    ])
    nxt_model = src_model.deep_clone()
    src_obs = tc.EVariable([nbatch, nobs], 0, "src_obs")
    nxt_obs = tc.EVariable([nbatch, nobs], 0, "nxt_obs")
    src_outmask = tc.EVariable([nbatch, nactions], 1, "src_outmask")
    nxt_outmask = tc.EVariable([nbatch], 1, "nxt_outmask")
    rewards = tc.EVariable([nbatch], 0, "rewards")

    def update(err, variables):
        half = len(variables) // 2
        src_vars, nxt_vars = variables[:half], variables[half:]
        src_updates = tc.api.approx.rms_momentum(err, src_vars, learning_rate=learning_rate, discount_factor=rms_discount,
                                                 apply=lambda x: tc.api.clip_by_l2norm(x, clip))
        assigns = []
        for nxt_var, (src_var, updated) in zip(nxt_vars, src_updates):
            diff = nxt_var - updated
            assigns.append((nxt_var, tc.api.assign_sub(nxt_var, target_update_rate * diff)))
        return assigns

    def error(models):
        src_act = models[0].connect(src_obs)
        nxt_act = models[1].connect(nxt_obs)
        target_vals = nxt_outmask * tc.api.reduce_max_1d(nxt_act, 0)
        future_reward = rewards + discount_rate * target_vals
        masked = tc.api.reduce_sum_1d(src_act * src_outmask, 0)
        return tc.api.reduce_mean(tc.api.square(masked - future_reward))

    train = tc.apply_update([src_model, nxt_model], update, error)
    # two forwards (source, target) + source backward without the input gradient
    flops = 2 * (2 * nbatch * nobs * nunits + 2 * nbatch * nunits * nactions) + 2 * nbatch * nobs * nunits + 4 * nbatch * nunits * nactions
    feeds = {"src_obs": src_obs, "nxt_obs": nxt_obs, "src_outmask": src_outmask, "nxt_outmask": nxt_outmask, "rewards": rewards}
    variables = src_model.get_storage() + nxt_model.get_storage()
    return Config(name, train, feeds, src_model, variables, flops,
                  "DQN %d-%d-%d x2 (source/target), rms_momentum + clip, replay batch %d" % (nobs, nunits, nactions, nbatch))


def dqn_batch(rng, cfg_feeds):
    """observations ~ U[0,1), one-hot action mask, rewards ~ U[-1,1]."""
    # shapes come back trimmed of trailing 1s (batch 1): recover the sizes from element counts
    nb = int(np.prod(cfg_feeds["rewards"].shape()))
    nobs = int(np.prod(cfg_feeds["src_obs"].shape())) // nb
    nact = int(np.prod(cfg_feeds["src_outmask"].shape())) // nb
    mask = np.zeros((nb, nact), dtype=np.float32)
    mask[np.arange(nb), rng.integers(0, nact, nb)] = 1
    return {
        "src_obs": rng.random((nb, nobs), dtype=np.float32),
        "nxt_obs": rng.random((nb, nobs), dtype=np.float32),
        "src_outmask": mask,
        "nxt_outmask": (rng.random(nb) < 0.95).astype(np.float32),
        "rewards": rng.uniform(-1, 1, nb).astype(np.float32),
    }


def cnn(in_ch=3, mid_ch=8, out_ch=4, width=12, height=10, nbatch=4, kernel_hw=(3, 3), learning_rate=0.5, seed=6, name="CNN"):
    """layer.conv2d -> sigmoid -> layer.conv2d -> sigmoid, MSE, SGD: the conv2d layer
    (cfg/tenncor/layer.yml:130-160,657-677 over nn.yml:48-98) driven like the gd_demo MLP. The reference
    ships no conv demo; this is the smallest model that exercises CONV forward,
    its kernel gradient (tensor-core GEMMs over a patch gather) and its image gradient."""
    tc.seed(seed)
    kh, kw = kernel_hw
    oh, ow = height - 2 * (kh - 1), width - 2 * (kw - 1)
    train_input = tc.EVariable([nbatch, height, width, in_ch], 0, "train_input")
    train_exout = tc.EVariable([nbatch, oh, ow, out_ch], 0, "train_exout")
    model = tc.api.layer.link([
        tc.api.layer.conv2d((kh, kw), in_ch, mid_ch),
        tc.api.layer.bind(tc.api.sigmoid),
        tc.api.layer.conv2d((kh, kw), mid_ch, out_ch),
        tc.api.layer.bind(tc.api.sigmoid),
    ], train_input)
    train = tc.apply_update(
        [model],
        lambda err, leaves: tc.api.approx.sgd(err, leaves, learning_rate=learning_rate),
        lambda models: tc.api.loss.mean_squared(train_exout, models[0].connect(train_input)))
    mh, mw = height - kh + 1, width - kw + 1
    # forward + kernel gradient of both layers, image gradient of the second; the post-update forward again
    flops = 2 * nbatch * kh * kw * (3 * mh * mw * in_ch * mid_ch + 4 * oh * ow * mid_ch * out_ch)
    return Config(name, train, {"x": train_input, "y": train_exout}, model, model.get_storage(), flops,
                  "CNN conv%dx%d %d-%d-%d sigmoid on %dx%d, MSE, SGD %.2g, batch %d" % (kh, kw, in_ch, mid_ch, out_ch, width, height, learning_rate, nbatch))


def conv_layer(in_ch=32, out_ch=64, width=34, height=34, nbatch=64, kernel_hw=(3, 3), learning_rate=0.5, seed=7, name="CONV"):
    """One layer.conv2d + sigmoid, MSE, SGD. No image gradient is built for the first layer's input
    (internal/teq/src/derive.cpp:151-159), so every CONV in the step lowers to patch gather + GEMM."""
    tc.seed(seed)
    kh, kw = kernel_hw
    oh, ow = height - kh + 1, width - kw + 1
    train_input = tc.EVariable([nbatch, height, width, in_ch], 0, "train_input")
    train_exout = tc.EVariable([nbatch, oh, ow, out_ch], 0, "train_exout")
    model = tc.api.layer.link([tc.api.layer.conv2d((kh, kw), in_ch, out_ch), tc.api.layer.bind(tc.api.sigmoid)], train_input)
    train = tc.apply_update(
        [model],
        lambda err, leaves: tc.api.approx.sgd(err, leaves, learning_rate=learning_rate),
        lambda models: tc.api.loss.mean_squared(train_exout, models[0].connect(train_input)))
    flops = 3 * 2 * nbatch * oh * ow * kh * kw * in_ch * out_ch  # forward, kernel gradient, post-update forward
    return Config(name, train, {"x": train_input, "y": train_exout}, model, model.get_storage(), flops,
                  "conv%dx%d %d->%d sigmoid on %dx%d, MSE, SGD %.2g, batch %d" % (kh, kw, in_ch, out_ch, width, height, learning_rate, nbatch))


def cnn_batch(rng, cfg_feeds):
    x = rng.random(cfg_feeds["x"].shape(), dtype=np.float32)
    y = rng.random(cfg_feeds["y"].shape(), dtype=np.float32)
    return x, y
