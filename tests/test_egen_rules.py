"""The opcode table's shape rules and redundancy rules, mirrored case by case from the reference's own
tests: internal/eigen/test/test_shaper.cpp (ShapeParser<OP>: expected shapes AND the exact error texts) and
internal/eigen/test/test_funcopt.cpp (FuncOpt<OP>: which functors make_funcattr elides). Host logic only."""
import re

import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def shape(opname, attrs, shapes):
    out = tc.egen.shape_parse(opname, attrs, shapes)
    while len(out) > 1 and out[-1] == 1:
        out.pop()
    return out


def fails(opname, attrs, shapes, message):
    with pytest.raises(Exception, match=re.escape(message)):
        tc.egen.shape_parse(opname, attrs, shapes)


NO_ARGS = "cannot operate without inputs"  # eigen::no_argument_err


# ---------------------------------------------------------------- test_shaper.cpp
def test_default():  # :20-45
    fails("ABS", {}, [], NO_ARGS)
    fails("ABS", {}, [[3, 4, 5], [3, 2, 5]], "cannot ABS with incompatible shapes [3\\2\\5\\1\\1\\1\\1\\1] and [3\\4\\5\\1\\1\\1\\1\\1]")
    assert shape("ABS", {}, [[3, 4, 6], [3, 4, 6]]) == [3, 4, 6]


def test_identity():  # :48-65
    fails("IDENTITY", {}, [], NO_ARGS)
    assert shape("IDENTITY", {}, [[3, 4, 5], [3, 2, 5]]) == [3, 4, 5]  # dependencies may have any shape


def test_reduce():  # :68-88
    attrs = {"rank_set": {3, 2, 5}}
    fails("REDUCE_SUM", attrs, [], NO_ARGS)
    assert shape("REDUCE_SUM", attrs, [[3, 4, 6, 7, 3]]) == [3, 4, 1, 1, 3]


def test_arg_reduce():  # :91-119
    fails("ARGMAX", {"rank": 3}, [], NO_ARGS)
    assert shape("ARGMAX", {"rank": 3}, [[3, 4, 6, 7, 3]]) == [3, 4, 6, 1, 3]
    assert shape("ARGMAX", {"rank": 8}, [[3, 4, 6, 7, 3]]) == [1]  # rank_cap: the flat index of the whole tensor


def test_permute():  # :122-137
    assert shape("PERMUTE", {"ranks": [2, 1, 0]}, [[3, 4, 6, 7, 3]]) == [6, 4, 3, 7, 3]


def test_extend():  # :140-166
    fails("EXTEND", {"dimensions": [1, 2, 0]}, [[3, 1, 6, 7, 3]], "cannot extend using zero dimensions [1\\2\\0]")
    fails("EXTEND", {"dimensions": [1, 2, 1]}, [[3, 4, 6, 7, 3]],
          "cannot extend non-singular dimension 1 of shape [3\\4\\6\\7\\3\\1\\1\\1]: bcast=[1\\2\\1]")
    assert shape("EXTEND", {"dimensions": [1, 2, 1]}, [[3, 1, 6, 7, 3]]) == [3, 2, 6, 7, 3]


def test_reshape():  # :169-195
    attrs = {"shape": [3, 4, 6]}
    fails("RESHAPE", attrs, [], NO_ARGS)
    fails("RESHAPE", attrs, [[3, 4, 5]], "cannot RESHAPE with shapes of different sizes 60 (shape [3\\4\\5\\1\\1\\1\\1\\1]) and 72 (shape [3\\4\\6\\1\\1\\1\\1\\1])")
    assert shape("RESHAPE", attrs, [[8, 3, 3]]) == [3, 4, 6]


def test_pad_slice_stride_scatter():  # :198-288
    pairs = {"dimension_pairs": [(3, 6), (2, 3), (0, 2)]}
    fails("PAD", pairs, [], NO_ARGS)
    assert shape("PAD", pairs, [[5, 4, 6, 7, 3]]) == [14, 9, 8, 7, 3]
    fails("SLICE", pairs, [], NO_ARGS)
    assert shape("SLICE", pairs, [[5, 4, 6, 7, 3]]) == [2, 2, 2, 7, 3]
    fails("STRIDE", {"dimensions": [3, 2, 3]}, [], NO_ARGS)
    assert shape("STRIDE", {"dimensions": [3, 2, 3]}, [[41, 4, 6, 7, 3]]) == [14, 2, 2, 7, 3]
    fails("SCATTER", {"shape": [3, 4, 6]}, [], NO_ARGS)
    assert shape("SCATTER", {"shape": [3, 4, 6]}, [[3, 4, 5]]) == [3, 4, 6]  # scatter allows a conflicting shape
    assert shape("SCATTER", {"shape": [3, 4, 6]}, [[8, 3, 3]]) == [3, 4, 6]


def test_contract():  # SHAPER.Matmul :291-350
    c = [3, 4, 6]
    assert shape("CONTRACT", {"rank_pairs": [(0, 0), (1, 1), (2, 2)]}, [c, c]) == [1]
    fails("CONTRACT", {"rank_pairs": [(0, 0), (0, 1), (2, 0)]}, [[2, 2, 2], [2, 2, 2]],
          "contraction dimensions [0:0\\0:1\\2:0] must be unique for each side")
    assert shape("CONTRACT", {"rank_pairs": []}, [c, c]) == [3, 4, 6, 3, 4, 6]  # outer product
    transposed = {"rank_pairs": [(1, 0)]}
    fails("CONTRACT", transposed, [c, c], "invalid shapes [3\\4\\6\\1\\1\\1\\1\\1] and [3\\4\\6\\1\\1\\1\\1\\1] do not match common dimensions [1:0]")
    assert shape("CONTRACT", transposed, [[4, 3, 6], [3, 5, 6]]) == [5, 6, 4, 6]
    assert shape("CONTRACT", transposed, [[4, 3], [3, 5]]) == [5, 4]
    typical = {"rank_pairs": [(0, 1)]}
    assert shape("CONTRACT", typical, [[3, 4, 6], [5, 3, 6]]) == [5, 6, 4, 6]
    assert shape("CONTRACT", typical, [[3, 4], [5, 3]]) == [5, 4]


def test_conv():  # :353-383
    img, kern = [4, 5, 6, 7], [3, 2, 5]
    fails("CONV", {"ranks": [0, 2]}, [img, kern],
          "cannot have ambiguous ranks not specified in kernelshape [3\\2\\5\\1\\1\\1\\1\\1] (ranks=[0\\2])")
    fails("CONV", {"ranks": [2, 1, 0]}, [img, kern],
          "cannot convolve a kernel of shape [3\\2\\5\\1\\1\\1\\1\\1] against smaller image of shape [4\\5\\6\\7\\1\\1\\1\\1] at dimensions (shape:kernel=0:2)")
    assert shape("CONV", {"ranks": [2, 1, 3]}, [img, kern]) == [4, 4, 4, 3]


def test_concat():  # :386-435
    a, b = [3, 4, 5, 6], [3, 4, 2, 6]
    fails("CONCAT", {"rank": 1}, [a, b], "cannot group concat incompatible shapes [3\\4\\5\\6\\1\\1\\1\\1] and [3\\4\\2\\6\\1\\1\\1\\1] along axis 1")
    assert shape("CONCAT", {"rank": 2}, [a, b]) == [3, 4, 7, 6]
    one, four = [3, 4, 1, 6], [3, 4, 4, 6]
    fails("CONCAT", {"rank": 1}, [one, one, four], "cannot group concat incompatible shapes [3\\4\\1\\6\\1\\1\\1\\1] and [3\\4\\4\\6\\1\\1\\1\\1] along axis 1")
    fails("CONCAT", {"rank": 2}, [one, four, one], "cannot group concat shapes with dimension that is not one")
    assert shape("CONCAT", {"rank": 2}, [one, one, one]) == [3, 4, 3, 6]


# ---------------------------------------------------------------- test_funcopt.cpp
@pytest.fixture
def a():
    return tc.variable(np.zeros((2, 2)), "a")  # teq shape [2, 2], DOUBLE


def opt(opname, attrs, args, out_dtype="DOUBLE"):
    return tc.egen.func_opt(opname, attrs, args, out_dtype)


def test_funcopt_default_and_add(a):  # :14-22, :221-232
    b, c = tc.variable(np.zeros((2, 2)), "b"), tc.variable(np.zeros((2, 2)), "c")
    assert not opt("SUB", {}, [a, b])
    assert opt("ADD", {}, [a]) and not opt("ADD", {}, [a, b]) and not opt("ADD", {}, [a, b, c])


def test_funcopt_reduce_argreduce(a):  # :24-62
    assert opt("REDUCE_SUM", {"rank_set": set()}, [a])
    assert not opt("REDUCE_SUM", {"rank_set": {1}}, [a])
    assert not opt("REDUCE_SUM", {"rank_set": {2}}, [a])  # the input's significant dimensions are not consulted
    assert not opt("ARGMAX", {"rank": 1}, [a])
    assert opt("ARGMAX", {"rank": 2}, [a])     # a singular rank: the index is always 0... the reference returns the argument
    assert not opt("ARGMAX", {"rank": 8}, [a])


def test_funcopt_permute_extend_reshape(a):  # :64-146
    assert not opt("PERMUTE", {"ranks": [1, 2, 0]}, [a]) and not opt("PERMUTE", {"ranks": [0, 2, 1]}, [a]) and not opt("PERMUTE", {"ranks": [1, 0]}, [a])
    assert opt("PERMUTE", {"ranks": []}, [a]) and opt("PERMUTE", {"ranks": [0, 1, 2]}, [a])
    assert opt("EXTEND", {}, [a]) and opt("EXTEND", {"dimensions": []}, [a]) and opt("EXTEND", {"dimensions": [1, 1, 1]}, [a])
    assert not opt("EXTEND", {"dimensions": [1, 1, 2]}, [a])
    assert not opt("RESHAPE", {"shape": [3, 2]}, [a]) and opt("RESHAPE", {"shape": [2, 2]}, [a])


def test_funcopt_slice_pad_cast(a):  # :148-256
    assert opt("SLICE", {"dimension_pairs": []}, [a])
    with pytest.raises(Exception, match=re.escape("cannot create slice with 0 dimensions (second value of extents) (extents=[1:2\\4:0])")):
        opt("SLICE", {"dimension_pairs": [(1, 2), (4, 0)]}, [a])
    assert opt("SLICE", {"dimension_pairs": [(0, 3), (0, 4)]}, [a])      # coverage beyond the shape
    assert not opt("SLICE", {"dimension_pairs": [(1, 2), (1, 3)]}, [a])
    assert opt("PAD", {"dimension_pairs": []}, [a]) and opt("PAD", {"dimension_pairs": [(0, 0), (0, 0)]}, [a])
    assert not opt("PAD", {"dimension_pairs": [(0, 3), (4, 0)]}, [a])
    assert opt("CAST", {}, [a], "DOUBLE") and not opt("CAST", {}, [a], "FLOAT") and not opt("CAST", {}, [a], "INT32")


def test_opcode_properties():
    """cfg/ops.yml `commutative` / `idempotent` flags (tools/egen/plugins/opcodes.py:96-141)"""
    ops = tc.egen.opcodes()
    assert len(ops) == 50 and ops[0] == "IDENTITY" and ops[-1] == "CAST"
    assert {o for o in ops if tc.egen.is_commutative(o)} == {"ADD", "MUL", "MIN", "MAX", "EQ", "NEQ"}
    assert {o for o in ops if not tc.egen.is_idempotent(o)} == {"RAND_UNIF", "ASSIGN_ADD", "ASSIGN_SUB", "ASSIGN_MUL", "ASSIGN_DIV", "CAST"}


# ---------------------------------------------------------------- test_typer.cpp (TypeParser<OP>)
def dtype(opname, attrs, dtypes):
    return tc.egen.type_parse(opname, attrs, dtypes)


def test_type_rules():  # internal/eigen/test/test_typer.cpp:13-87
    with pytest.raises(Exception, match=NO_ARGS):
        dtype("ADD", {}, [])
    # the result has the highest precision among the arguments, whatever their order
    assert dtype("ADD", {}, ["DOUBLE", "FLOAT"]) == "DOUBLE" and dtype("ADD", {}, ["FLOAT", "DOUBLE"]) == "DOUBLE"
    assert dtype("ADD", {}, ["FLOAT", "INT32"]) == "FLOAT" and dtype("ADD", {}, ["INT32", "FLOAT"]) == "FLOAT"
    # ASSIGN takes the destination's type
    with pytest.raises(Exception, match=NO_ARGS):
        dtype("ASSIGN", {}, [])
    assert dtype("ASSIGN", {}, ["DOUBLE", "FLOAT"]) == "DOUBLE" and dtype("ASSIGN", {}, ["FLOAT", "DOUBLE"]) == "FLOAT"
    assert dtype("ASSIGN", {}, ["INT32", "DOUBLE"]) == "INT32" and dtype("ASSIGN", {}, ["INT32", "FLOAT"]) == "INT32"
    # CAST: identity without the dtype attribute, the attribute's type with it
    with pytest.raises(Exception, match=NO_ARGS):
        dtype("CAST", {}, [])
    assert [dtype("CAST", {}, [t]) for t in ("DOUBLE", "FLOAT", "INT32")] == ["DOUBLE", "FLOAT", "INT32"]
    assert [dtype("CAST", {"dtype": "INT32"}, [t]) for t in ("DOUBLE", "FLOAT", "INT32")] == ["INT32"] * 3


# ---------------------------------------------------------------- test_packer.cpp (attribute packers' error behaviour)
def test_attribute_packers():  # internal/eigen/test/test_packer.cpp:116-300
    """a rule that needs an attribute names the missing key; packers refuse ranks beyond rank_cap with the reference's texts"""
    for opname, key in [("REDUCE_SUM", "rank_set"), ("PERMUTE", "ranks"), ("CONTRACT", "rank_pairs"), ("PAD", "dimension_pairs"),
                        ("ARGMAX", "rank"), ("RESHAPE", "shape"), ("STRIDE", "dimensions")]:
        shapes = [[3, 4], [3, 4]] if opname == "CONTRACT" else [[3, 4]]
        fails(opname, {}, shapes, "cannot find `%s` attribute" % key)
    fails("PERMUTE", {"ranks": [8, 3, 4, 10]}, [[3, 4]], "cannot reference ranks beyond rank_cap 8: [8\\3\\4\\10]")
    fails("CONTRACT", {"rank_pairs": [(8, 3), (4, 10)]}, [[3, 4], [3, 4]], "cannot reference ranks beyond rank_cap 8: [8:3\\4:10]")
    fails("REDUCE_SUM", {"rank_set": {3, 4, 8, 10}}, [[3, 4]], "cannot reference ranks beyond rank_cap 8: [3\\4\\8\\10]")
    fails("REDUCE_SUM", {"rank_set": set(range(9))}, [[3, 4]], "cannot specify 9 ranks when 8 (rank_cap) are available")
    # values survive the round trip through the attribute map
    assert shape("PAD", {"dimension_pairs": [(2, 1), (0, 3)]}, [[3, 4]]) == [6, 7]
    assert shape("REDUCE_SUM", {"rank_set": {1, 8}}, [[3, 4]]) == [3]  # rank_cap itself is accepted and names no rank
