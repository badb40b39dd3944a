"""The reference's own demo scripts, unmodified, against this package ("the python demos run unchanged with only the
device changing"). Runs only where the reference checkout exists (this container, never the GPU box) and no GPU does: each script must get
through imports, argument parsing, data loading, model construction, derivative and update graphs — every host-side API call it makes — and
stop exactly where it first asks the device for numbers, with the back end's "no CUDA device" error (there is no CPU fallback to carry on with)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMOS = ["gd_demo.py", "dqn_demo.py", "dbn_demo.py", "lstm/fast_demo.py", "gru/fast_demo.py", "lstm/latin_demo.py", "gru/latin_demo.py"]

RUNNER = """
import runpy, sys
import tenncor_b200.compat as compat
compat.install()
sys.argv = [sys.argv[1]]
runpy.run_path(sys.argv[0], run_name="__main__")
"""


def _has_gpu():
    try:
        from tenncor_b200 import cabi
        return cabi.lib().tcr_device_count() > 0
    except Exception:
        return False


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "demo")), reason="reference checkout not present")
@pytest.mark.parametrize("demo", DEMOS)
def test_reference_demo_reaches_the_device(built, demo, tmp_path):
    if _has_gpu():
        pytest.skip("a device is present: the demo would train for minutes")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), HOME=str(tmp_path), TMPDIR=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", RUNNER, os.path.join(REF, "demo", demo)], cwd=REF, env=env, capture_output=True, text=True, timeout=300)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode != 0, tail
    assert "no CUDA device visible (this back end has no CPU fallback)" in r.stderr, tail
    last_frame = [line for line in r.stderr.split("\n") if line.strip().startswith(("File ", "p = ", "err", "train"))]
    assert "AttributeError" not in r.stderr and "TypeError" not in r.stderr and "ImportError" not in r.stderr, tail
    assert last_frame, tail
