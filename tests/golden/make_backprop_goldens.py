"""Extracts the expected derivative graphs of tenncor/eteq/test/test_backprop.cpp (EXPECT_GRAPHEQ string literals, one per
TEST(BACKPROP, Name)) into tests/golden/backprop_goldens.json. Only the literals are copied; tests/test_backprop_golden.py
rebuilds each operand set through our own host and prints our graph in the same PrettyEquation format.
Run in the build container:  python tests/golden/make_backprop_goldens.py"""
import json
import os
import re

REF = "/root/reference/tenncor/eteq/test/test_backprop.cpp"
# the optimizer update graphs asserted the same way (EXPECT_GRAPHEQ) by tenncor/test/test_approx.cpp
APPROX_REF = "/root/reference/tenncor/test/test_approx.cpp"
LAYER_REF = "/root/reference/tenncor/test/test_layer.cpp"


def extract(path, suite, out, prefix=""):
    src = open(path).read()
    for m in re.finditer(r"TEST\(%s, (\w+)\)\n\{(.*?)\n\}\n" % suite, src, re.S):
        name, body = prefix + m.group(1), m.group(2)
        graphs = []
        # runs of adjacent C string literals (only whitespace between them); a run that starts with "(" is one expected graph
        lits = [(l.start(), l.end(), l.group(1)) for l in re.finditer(r'"((?:[^"\\]|\\.)*)"', body)]
        run = []
        for i, (a, b, text) in enumerate(lits):
            if run and body[lits[i - 1][1]:a].strip() == "":
                run.append(text)
            else:
                if run:
                    graphs.append(run)
                run = [text]
        if run:
            graphs.append(run)
        graphs = ["".join(r).replace("\\\\", "\\").replace("\\n", "\n") for r in graphs]
        graphs = [g for g in graphs if g.startswith("(") and "\n" in g]
        if graphs:
            line = src[:m.start()].count("\n") + 1
            out[name] = {"cite": "%s:%d" % (path.replace("/root/reference/", ""), line), "graphs": graphs}


def main():
    out = {}
    extract(REF, "BACKPROP", out)
    extract(APPROX_REF, "APPROX", out, prefix="Approx")
    # layer connection graphs (tenncor/test/test_layer.cpp:36-262): dense / conv2d / rbm (both directions) / bind
    for suite in ("DENSE", "CONV", "RBM", "BIND"):
        extract(LAYER_REF, suite, out, prefix="Layer" + suite.capitalize())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "backprop_goldens.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print({k: len(v["graphs"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
