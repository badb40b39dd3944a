"""ONNX-dialect model files: tc.load_from_file / tc.save_to_file mirror
tenncor/python/eteq_ext.cpp:408-487 over internal/onnx + tenncor/serial.

Pinned against the reference: tests/golden/reference_models.json holds files written by the reference's
serializer (its shipped demo models gd / dqn / dbn / rnn). Each is decoded twice — by the host's loader and by the small independent
wire-format reader below — and the forward pass the loaded graph defines (evaluated by the CPU oracle)
must equal the one recomputed in numpy from the independently decoded weights. No device needed."""
import os
import struct

import numpy as np
import pytest

import tenncor_b200 as tc
from oracle import tcr_oracle as orc
from tenncor_b200 import configs

GOLDEN_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_models.json")


def golden_model(name, tmp_dir=None):
    """path of the reference's models/<name>.onnx, materialised from the committed fixture (tests/golden/make_model_goldens.py)"""
    import base64
    import hashlib
    import json
    import tempfile
    entry = json.load(open(GOLDEN_JSON))["models"][name]
    data = base64.b64decode(entry["base64"])
    assert len(data) == entry["bytes"] and hashlib.sha256(data).hexdigest() == entry["sha256"]
    d = tmp_dir or tempfile.mkdtemp(prefix="tcr_golden_")
    path = os.path.join(str(d), name + ".onnx")
    with open(path, "wb") as f:
        f.write(data)
    return path


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


# ---------------------------------------------------------------- independent protobuf reader (test side only)
def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return r, i


def _fields(b):
    i, out = 0, []
    while i < len(b):
        k, i = _varint(b, i)
        f, w = k >> 3, k & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(w)
        out.append((f, w, v))
    return out


def _graph_of(model_bytes):
    return [v for f, w, v in _fields(model_bytes) if f == 7][0]


def _initializers(graph_bytes, out=None, labels=None):
    """every float initializer of the graph and its nested layer graphs: id -> array, id -> label"""
    out = {} if out is None else out
    labels = {} if labels is None else labels
    for f, w, v in _fields(graph_bytes):
        if f == 5:
            t = _fields(v)
            name = [x for ff, ww, x in t if ff == 8][0].decode()
            dims = []
            for ff, ww, x in t:
                if ff == 1:
                    dims += list(x) if ww == 2 else [x]
            data = b"".join(x for ff, ww, x in t if ff == 4)
            out[name] = (dims, np.frombuffer(data, "<f4").astype(np.float64))
        elif f == 14:
            a = _fields(v)
            tid = [x for ff, ww, x in a if ff == 1][0].decode()
            for ff, ww, x in a:
                if ff == 2:
                    kv = {k: val.decode() for k, _, val in _fields(x)}
                    if kv.get(1) == "TENSOR_NAME":
                        labels[tid] = kv.get(2, "")
        elif f == 1:
            for ff, ww, x in _fields(v):
                if ff == 5:  # attribute
                    for f3, w3, x3 in _fields(x):
                        if f3 == 6:
                            _initializers(x3, out, labels)
    return out, labels


def _oracle_eval(roots):
    tape = tc.dump_graph(roots)
    ids = tc.dump_ids(roots, tape)
    vals = orc.eval_tape(tape)
    return [np.asarray(vals[ids[r]], np.float64).reshape(-1) for r in roots]


def _sigmoid(x):
    return 1 / (1 + np.exp(-x))


# ---------------------------------------------------------------- the reference's own files
@pytest.mark.parametrize("name,widths", [("gd", [10, 9, 5]), ("dqn", [10, 9, 9])])
def test_reference_dense_models_load_and_evaluate(name, widths):
    path = golden_model(name)
    model = tc.load_from_file(path)
    assert len(model) == 1
    got = _oracle_eval(model)[0]
    inits, labels = _initializers(_graph_of(open(path, "rb").read()))
    weights = [v for k, v in inits.items() if labels.get(k) == "weight"]
    biases = [v for k, v in inits.items() if labels.get(k) == "bias"]
    inputs = [v for k, v in inits.items() if labels.get(k) not in ("weight", "bias")]
    assert len(weights) == 2 and len(biases) == 2 and len(inputs) == 1
    (xd, x) = inputs[0]
    assert xd[0] == widths[0]
    h = x.reshape(xd[1], xd[0])  # teq [in, B] = row-major B x in
    for (wd, w), (bd, b) in zip(weights, biases):
        assert wd[1] == h.shape[1]
        h = _sigmoid(h @ w.reshape(wd[1], wd[0]) + b[None, :])
    np.testing.assert_allclose(got, h.reshape(-1), rtol=1e-6)  # the file holds fp32 weights; both sides compute in fp32 / double
    assert len(model[0].get_storage()) == 4  # weight + bias of both dense layers came back as variables


@pytest.mark.parametrize("name", ["gd", "dqn", "dbn", "rnn"])
def test_reference_files_survive_a_round_trip(name, tmp_path):
    path = golden_model(name)
    first = tc.load_from_file(path)
    want = _oracle_eval(first)
    again_path = str(tmp_path / (name + "_again.onnx"))
    assert tc.save_to_file(again_path, first)
    second = tc.load_from_file(again_path)
    assert len(second) == len(first)
    for w, g in zip(want, _oracle_eval(second)):
        np.testing.assert_array_equal(g, w)
    assert [len(m.get_storage()) for m in second] == [len(m.get_storage()) for m in first]


# ---------------------------------------------------------------- models built here
def test_layer_structure_and_values_survive_save_and_load(tmp_path):
    cfg = configs.mlp(10, 9, 5, 3)
    rng = np.random.default_rng(0)
    x, _ = configs.mlp_batch(rng, cfg.feeds)
    cfg.feeds["x"].assign(x)
    want = _oracle_eval([cfg.model])[0]
    path = str(tmp_path / "mlp.onnx")
    assert tc.save_to_file(path, [cfg.model])
    loaded = tc.load_from_file(path)[0]
    np.testing.assert_array_equal(_oracle_eval([loaded])[0], want)
    # the nested "layer" graphs were rebuilt: the loaded model re-connects to a new input and shares its variables
    assert loaded.opname() == "IDENTITY" and len(loaded.get_storage()) == 4
    other = tc.variable(rng.random((7, 10), dtype=np.float32), "other")
    y = loaded.connect(other)
    assert y.shape() == [7, 5]
    # the file itself follows the reference's layout: one _LINK node whose "layer" attribute nests the sub-layers
    g = _graph_of(open(path, "rb").read())
    nodes = [v for f, w, v in _fields(g) if f == 1]
    assert len(nodes) == 1
    node = _fields(nodes[0])
    assert [x for f, w, x in node if f == 4][0] == b"_LINK"
    attr = _fields([x for f, w, x in node if f == 5][0])
    assert [x for f, w, x in attr if f == 1][0] == b"layer" and [x for f, w, x in attr if f == 20][0] == 5  # GRAPH
    sub_ops = [[y for ff, ww, y in _fields(x) if ff == 4][0] for f, w, x in _fields([x for f, w, x in attr if f == 6][0]) if f == 1]
    assert sub_ops == [b"_DENSE_LAYER", b"_UNARY_BIND", b"_DENSE_LAYER", b"_UNARY_BIND", b"IDENTITY"]


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_recurrent_model_round_trip(kind, tmp_path):
    cfg = configs.recurrent(kind, vocab=6, hidden=5, seq=4, batch=None)
    rng = np.random.default_rng(1)
    x, _ = configs.recurrent_batch(rng, cfg.feeds, 6)
    probe = tc.variable(x, "probe")
    want = _oracle_eval([cfg.model.connect(probe)])[0]
    path = str(tmp_path / (kind + ".onnx"))
    assert tc.save_to_file(path, [cfg.model])
    loaded = tc.load_from_file(path)[0]
    np.testing.assert_array_equal(_oracle_eval([loaded.connect(probe)])[0], want)
    assert len(loaded.get_storage()) == len(cfg.model.get_storage())


def test_keys_and_precedence(tmp_path):
    a = tc.variable(np.arange(6, dtype=np.float32).reshape(2, 3), "a")
    b = tc.variable(np.ones((2, 3), dtype=np.float32), "b")
    s, p = a + b, a * b
    path = str(tmp_path / "two.onnx")
    assert tc.save_to_file(path, [s, p], keys={"sum": s, "prod": p})
    first = tc.load_from_file(path, key_prec={"prod": 0, "sum": 1})
    assert [t.opname() for t in first] == ["MUL", "ADD"]
    second = tc.load_from_file(path, key_prec={"sum": 0, "prod": 1})
    assert [t.opname() for t in second] == ["ADD", "MUL"]
    np.testing.assert_array_equal(_oracle_eval(second)[0], (np.arange(6) + 1).astype(np.float64))


def test_dtypes_attributes_and_placeholders(tmp_path):
    rng = np.random.default_rng(2)
    x = tc.variable(rng.integers(-5, 5, (3, 4, 5)).astype(np.int32), "x")
    d = tc.variable(rng.random((3, 4, 5)), "d")  # float64
    roots = [tc.api.permute(x, [2, 0, 1]), tc.api.reduce_sum(d, 1, 1), tc.api.slice(x, 1, 2, 1), tc.api.pad(d, (1, 2), 0),
             tc.api.contract(d, d, [(0, 0), (1, 1)]), tc.api.extend(tc.api.reduce_max(x, 0, 1), 0, [3])]
    want = _oracle_eval(roots)
    path = str(tmp_path / "ops.onnx")
    assert tc.save_to_file(path, roots)
    loaded = tc.load_from_file(path)
    assert [t.opname() for t in loaded] == [t.opname() for t in roots]
    assert [str(t.dtype()) for t in loaded] == [str(t.dtype()) for t in roots]
    for w, g in zip(want, _oracle_eval(loaded)):
        np.testing.assert_array_equal(g, w)


def test_errors():
    with pytest.raises(Exception, match="not found"):
        tc.load_from_file("/nonexistent/model.onnx")
    assert tc.save_to_file("/tmp/_tcr_never_written.onnx", []) is False


def test_pretrained_gd_model_solves_the_gd_demo_task():
    """models/gd.onnx is the trained gd_demo MLP (demo/gd_demo.py:76,99-118 'pretrained mlp error rate'): connected
    to fresh inputs it must reproduce y = pairwise mean of x far better than an untrained model of the same shape."""
    rng = np.random.default_rng(0)
    x = rng.random((200, 10)).astype(np.float32)
    y = (x[:, 0::2] + x[:, 1::2]) / 2
    testin = tc.variable(x, "testin")
    pretrained = tc.load_from_file(golden_model("gd"))[0].connect(testin)
    untrained = configs.mlp(10, 9, 5, 3).model.connect(testin)
    err = [float(np.mean(np.abs(o.reshape(200, 5) - y))) for o in _oracle_eval([pretrained, untrained])]
    assert err[0] < 0.05 and err[1] > 3 * err[0], err


def test_conv_and_rbm_models_round_trip(tmp_path):
    """the other layer kinds the demos save: a conv2d stack (_CONV_LAYER) and an RBM's two halves saved together under keys and
    rebuilt as tc.RBMLayer(*tc.load_from_file(...)) (demo/rbm_demo.py:75-78,177-180)"""
    cfg = configs.cnn()
    rng = np.random.default_rng(4)
    probe = tc.variable(rng.random(cfg.feeds["x"].shape(), dtype=np.float32), "probe")
    want = _oracle_eval([cfg.model.connect(probe)])[0]
    path = str(tmp_path / "cnn.onnx")
    assert tc.save_to_file(path, [cfg.model])
    loaded = tc.load_from_file(path)[0]
    np.testing.assert_array_equal(_oracle_eval([loaded.connect(probe)])[0], want)
    assert len(loaded.get_storage()) == 4  # two kernels, two biases
    # the reloaded conv2d still lowers to the tensor-core form: the composite's structure survived the file
    assert sum(s.startswith("CONV2D im2col+GEMM") for s in tc.describe_plan([loaded.connect(probe)])) == 2

    rbm = tc.api.layer.rbm(12, 5)
    vis = tc.variable((rng.random((3, 12)) < 0.5).astype(np.float32), "vis")
    hid = tc.variable((rng.random((3, 5)) < 0.5).astype(np.float32), "hid")
    want_f, want_b = _oracle_eval([rbm.connect(vis), rbm.backward_connect(hid)])
    path = str(tmp_path / "rbm.onnx")
    assert tc.save_to_file(path, [rbm.fwd(), rbm.bwd()], keys={"fwd": rbm.fwd(), "bwd": rbm.bwd()})
    again = tc.RBMLayer(*tc.load_from_file(path, key_prec={"fwd": 0, "bwd": 1}))
    got_f, got_b = _oracle_eval([again.connect(vis), again.backward_connect(hid)])
    np.testing.assert_array_equal(got_f, want_f)
    np.testing.assert_array_equal(got_b, want_b)
    # the weight is ONE variable shared by both halves, in the file as in memory
    assert len(again.fwd().get_storage()) == len(rbm.fwd().get_storage())
    weight = [v for v in again.fwd().get_storage() if v.shape() == [12, 5]][0]
    weight.assign(np.zeros((12, 5), dtype=np.float32))
    np.testing.assert_array_equal(_oracle_eval([again.backward_connect(hid)])[0], np.zeros(3 * 12))  # bwd sees fwd's zeroed weight (vbias starts at 0)
