"""host/egen.{hpp,cpp} restates by hand what the reference GENERATES from cfg/ops.yml and its type file
(tools/egen/plugins/opcodes.py:143-187, dtypes.py:8-145). This test re-applies the generator's own rules to those
files and diffs the result against the compiled tables, so the hand-written surface cannot drift from the config:
  _GENERATED_OPCODE     = the keys of opcode.opcalls in file order, numbered from 1 (BAD_OP = 0)
  is_commutative(op)    = opcalls[op].get("commutative", False)
  is_idempotent(op)     = opcalls[op].get("idempotent", True)
  _GENERATED_DTYPE      = the keys of `dtype` in file order; ctype size, `precision`, default_type
The reference tree exists only in the build container: skipped elsewhere (the tables are also pinned value by value in
tests/test_egen_rules.py)."""
import ctypes
import os

import pytest

import tenncor_b200 as tc

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "cfg", "ops.yml")), reason="reference tree not present")

CTYPE_BYTES = {"double": 8, "float": 4, "int8_t": 1, "uint8_t": 1, "int16_t": 2, "uint16_t": 2, "int32_t": 4, "uint32_t": 4,
               "int64_t": 8, "uint64_t": 8}


def _load(name):
    import yaml
    with open(os.path.join(REF, "cfg", name)) as f:
        return yaml.safe_load(f)


def test_opcode_enum_and_flags_match_the_generator_rules():
    opcalls = _load("ops.yml")["opcode"]["opcalls"]
    want = list(opcalls.keys())
    assert tc.egen.opcodes() == want
    for name, call in opcalls.items():
        assert tc.egen.is_commutative(name) == bool(call.get("commutative", False)), name
        assert tc.egen.is_idempotent(name) == bool(call.get("idempotent", True)), name


def test_dtype_enum_matches_the_full_type_file():
    spec = _load("fulltype.yml")
    want = [(name, CTYPE_BYTES[d["ctype"]], int(d["precision"])) for name, d in spec["dtype"].items()]
    assert [tuple(t) for t in tc.egen.dtypes()] == want
    assert tc.egen.default_dtype() == spec["default_type"]
    for name, d in spec["dtype"].items():
        assert CTYPE_BYTES[d["ctype"]] == ctypes.sizeof(getattr(ctypes, "c_" + d["ctype"].replace("_t", ""))) if d["ctype"] not in ("double", "float") else True
