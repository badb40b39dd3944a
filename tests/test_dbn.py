"""Deep belief network trainer (tenncor/trainer/dbn.hpp, demo/dbn_demo.py; BASELINE config 2 "RBM / DBN contrastive-divergence
training"). Graph construction is host logic (CPU); training follows the demo on the GPU: RAND_UNIF streams differ from the
reference's std::default_random_engine by design, so the demo's outcome is checked, not its numbers."""
import numpy as np
import pytest

import tenncor_b200 as tc

X = np.array([[1, 1, 1, 0, 0, 0], [1, 0, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [0, 0, 1, 1, 1, 0],
              [0, 0, 1, 1, 0, 0], [0, 0, 1, 1, 1, 0], [0, 0, 1, 1, 0, 1]], dtype=np.float32)  # demo/dbn_demo.py:46-54
Y = np.array([[1, 0], [1, 0], [1, 0], [0, 1], [0, 1], [0, 1], [0, 0]], dtype=np.float32)


def build(cdk=1):
    tc.seed(123)
    rbms = [tc.api.layer.rbm(6, 3), tc.api.layer.rbm(3, 3)]
    dense = tc.api.layer.dense([3], [2], kernel_init=tc.api.init.zeros())
    inter = [e for rbm in rbms for e in (rbm.fwd(), tc.api.layer.bind(tc.api.sigmoid))]
    model = tc.api.layer.link(inter + [dense, tc.api.layer.bind(lambda x: tc.api.softmax(x, 0, 1))])
    trainer = tc.DBNTrainer(rbms, dense, 0, X.shape[0], pretrain_lr=0.1, train_lr=0.1, cdk=cdk)
    return rbms, dense, model, trainer


def test_trainer_graphs(built):
    tc.require_host()
    rbms, dense, model, trainer = build(cdk=3)
    pipes = trainer.sample_pipes()
    assert [p.shape() for p in pipes] == [[7, 6], [7, 3], [7, 3]]  # trainx, then one hidden sample per rbm
    updates = trainer.update_graphs()
    assert [u.opname() for u in updates] == ["ASSIGN_ADD"] * 6 + ["ASSIGN"]  # (weight, hbias, vbias) x 2 layers, then the learning-rate decay
    # each layer's CD-3 chain: the given hidden sample plus two Gibbs round trips = 4 more bernoulli draws, all under the weight update
    ops = [n["op"] for n in tc.dump_graph([updates[0]]) if n["kind"] != "leaf"]
    assert ops.count("RAND_UNIF") == 1 + 2 * 2
    tail = tc.dump_graph([updates[-1]])
    assert sum(n["op"] == "ASSIGN_ADD" for n in tail if n["kind"] != "leaf") == 2  # dense weight and bias ride on the decay's IDENTITY
    labels = {n.get("label") for n in tail if n["kind"] == "leaf"}
    assert {"learning_rate", "trainy", "weight", "bias"} <= labels
    # the RBM layer wrapper round-trips through its two halves (tc.RBMLayer(fwd, bwd), demo/rbm_demo.py:75)
    again = tc.RBMLayer(rbms[0].fwd(), rbms[0].bwd())
    assert again.connect(pipes[0]).shape() == [7, 3] and again.backward_connect(pipes[1]).shape() == [7, 6]


@pytest.fixture
def evaluator(request):
    tc.set_evaluator(request.param)
    yield request.param
    tc.set_evaluator("plan")


@pytest.mark.gpu
@pytest.mark.parametrize("evaluator", ["node", "plan"], indirect=True)
def test_dbn_demo_learns_its_toy_problem(gpu, evaluator):
    rbms, dense, model, trainer = build(cdk=1)
    seen = []
    c0 = [trainer.reconstruction_cost(0), trainer.reconstruction_cost(1)]
    trainer.pretrain(X, nepochs=1000, logger=lambda epoch, layer: seen.append((layer, epoch)))  # the demo's defaults: 1000 / 200 epochs, CD-1
    assert seen[0] == (0, 0) and seen[-1] == (1, 999) and len(seen) == 2000
    c1 = [trainer.reconstruction_cost(0), trainer.reconstruction_cost(1)]
    assert np.isfinite(c0 + c1).all() and c1[0] < c0[0], (c0, c1)  # the first RBM reconstructs its input better than at the start
    t0 = trainer.training_cost()
    trainer.finetune(X, Y, nepochs=200)
    t1 = trainer.training_cost()
    assert np.isfinite([t0, t1]).all() and t1 < t0, (t0, t1)
    # demo/dbn_demo.py:100-112: an input like the first three rows is classified like them
    probe = tc.variable(np.array([1, 1, 0, 0, 0, 0], dtype=np.float32), "probe")
    out = model.connect(probe).get().reshape(-1)
    assert out[0] > out[1], out
    probe2 = tc.variable(np.array([0, 0, 0, 1, 1, 0], dtype=np.float32), "probe2")
    out2 = model.connect(probe2).get().reshape(-1)
    assert out2[1] > out2[0], out2
