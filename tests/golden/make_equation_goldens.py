"""Extracts the golden vectors of the reference's end-to-end equation tests into
tests/golden/equation_goldens.json.

Source: /root/reference/tenncor/test/test_equation.cpp — matmul_complex (:26-136),
contract_equivalent (:139-250), sigmoid_MLP_slow (:252-480), sigmoid_MLP_fast (:482-707),
tanh_RNN (:709-840), tanh_RNN_layer (:842-970). The reference asserts these with
EXPECT_DOUBLE_EQ (4 ulp) against its Eigen back end. Only the literal data
(`std::vector<double> x = {...}`, `teq::Shape s({..})`, `teq::DimsT l = {..}`) is copied;
the graphs are rebuilt through our own API in tests/test_equation_golden.py.
Run here (the reference tree does not exist on the GPU box):
    python tests/golden/make_equation_goldens.py
"""
import json
import os
import re

REF = "/root/reference/tenncor/test/test_equation.cpp"
# one more golden of the same kind from the API tests: forward and both gradients of CONV
# (tenncor/test/test_api.cpp:2265-2361, asserted there with ASSERT_VECEQ on doubles)
API_REF = "/root/reference/tenncor/test/test_api.cpp"
API_FUNCS = {"api_convolution": (2265, 2361)}
# layer-level goldens: rnn / dense stacks connected through layer.link, full gradients (tenncor/test/test_layer.cpp CONNECT.*)
LAYER_REF = "/root/reference/tenncor/test/test_layer.cpp"
LAYER_FUNCS = {"layer_tanh_rnn": (264, 388), "layer_dense_tanh_rnn": (390, 572), "layer_tanh_rnn_full": (574, 814),
               "layer_tanh_rnn_cross_entropy": (816, 1059)}
FUNCS = {"matmul_complex": (26, 136), "contract_equivalent": (139, 250), "sigmoid_MLP_slow": (252, 480),
         "sigmoid_MLP_fast": (482, 707), "tanh_RNN": (709, 840), "tanh_RNN_layer": (842, 970)}


def numbers(body):
    return [float(tok) for tok in re.findall(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)", body)]


def main():
    lines = open(REF).read().split("\n")
    out = {}
    for name, (lo, hi) in FUNCS.items():
        text = "\n".join(lines[lo - 1:hi])
        vecs = {m.group(1): numbers(m.group(2)) for m in re.finditer(r"std::vector<double>\s+(\w+)\s*=\s*\{([^}]*)\};", text, re.S)}
        shapes = {m.group(1): [int(v) for v in numbers(m.group(2))] for m in re.finditer(r"teq::Shape\s+(\w+)\(\{([^}]*)\}\)", text)}
        shapes.update({m.group(1): [int(v) for v in numbers(m.group(2))] for m in re.finditer(r"teq::DimsT\s+(\w+)\s*=\s*\{([^}]*)\};", text)})
        out[name] = {"cite": "tenncor/test/test_equation.cpp:%d-%d" % (lo, hi), "vectors": vecs, "shapes": shapes}
    api_lines = open(API_REF).read().split("\n")
    for name, (lo, hi) in API_FUNCS.items():
        text = "\n".join(api_lines[lo - 1:hi])
        vecs = {m.group(1): numbers(m.group(2)) for m in re.finditer(r"std::vector<double>\s+(\w+)\s*=\s*\{([^}]*)\};", text, re.S)}
        shapes = {m.group(1): [int(v) for v in numbers(m.group(2))] for m in re.finditer(r"teq::DimsT\s+(\w+)\s*=\s*\{([^}]*)\};", text, re.S)}
        out[name] = {"cite": "tenncor/test/test_api.cpp:%d-%d" % (lo, hi), "vectors": vecs, "shapes": shapes}
    layer_lines = open(LAYER_REF).read().split("\n")
    for name, (lo, hi) in LAYER_FUNCS.items():
        text = "\n".join(layer_lines[lo - 1:hi])
        vecs = {m.group(1): numbers(m.group(2)) for m in re.finditer(r"std::vector<double>\s+(\w+)\s*=\s*\{([^}]*)\};", text, re.S)}
        dims = {m.group(1): int(m.group(2)) for m in re.finditer(r"teq::(?:DimT|RankT)\s+(\w+)\s*=\s*(\d+);", text)}
        out[name] = {"cite": "tenncor/test/test_layer.cpp:%d-%d" % (lo, hi), "vectors": vecs, "shapes": {}, "dims": dims}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "equation_goldens.json")
    with open(path, "w") as f:
        json.dump(out, f)
    for k, v in out.items():
        print(k, {n: len(x) for n, x in v["vectors"].items()}, v["shapes"])


if __name__ == "__main__":
    main()
