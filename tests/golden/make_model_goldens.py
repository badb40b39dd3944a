"""Collect the reference's shipped demo models (models/*.onnx, written by the reference's own serializer) into one
fixture. Run in the build container, where /root/reference exists:

    python tests/golden/make_model_goldens.py

Only the small files are taken (gd, dqn, dbn, rnn: 34 KB together); they are data the loader is pinned against
(tests/test_onnx.py), stored base64-encoded with their byte length and sha256."""
import base64
import hashlib
import json
import os

REF = "/root/reference/models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_models.json")
NAMES = ["gd", "dqn", "dbn", "rnn"]
# models/test/<name>.onnx + <name>.txt: the file and the PrettyEquation rendering the reference's own serializer tests compare
# against (tenncor/test/test_serialize.cpp and tenncor/serial/test/test_serialize.cpp, SaveGraph / LoadGraph)
TEST_NAMES = ["eteq", "serial"]


def main():
    out = {"_source": "mingkaic/tenncor models/<name>.onnx (ONNX-dialect ModelProto written by tcr::save_model)", "models": {}}
    for name in NAMES:
        data = open(os.path.join(REF, name + ".onnx"), "rb").read()
        out["models"][name] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(), "base64": base64.b64encode(data).decode()}
    out["test_models"] = {}
    for name in TEST_NAMES:
        data = open(os.path.join(REF, "test", name + ".onnx"), "rb").read()
        out["test_models"][name] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(), "base64": base64.b64encode(data).decode(),
                                    "txt": open(os.path.join(REF, "test", name + ".txt")).read()}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
