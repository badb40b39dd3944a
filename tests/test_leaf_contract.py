"""Leaf storage, metadata, the subscriber set and the TypeCaster, mirrored from
tenncor/eteq/test/test_variable.cpp, test_constant.cpp, test_caster.cpp and internal/eigen/test/test_meta.cpp,
test_observable.cpp. Leaves stage their data on the host until a device first needs it, so all of this runs without a GPU."""
import gc

import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


BIG = np.arange(1, 13).reshape(4, 3)  # teq::Shape({3, 4}); numpy lists the slow dimension first


def trio():
    return (tc.variable(BIG.astype(np.float64), "A"), tc.variable(BIG.astype(np.float32), "B"),
            tc.variable(BIG.astype(np.int32), "C"))


def test_variable_copy():  # VARIABLE.CopyMove :15-46 (move construction has no counterpart above the pybind boundary)
    for var, label in zip(trio(), "ABC"):
        assert str(var) == label
        cpy = var.clone()
        assert str(cpy) == label and cpy != var
        np.testing.assert_array_equal(cpy.data(), var.data())
        tc.to_variable(var).assign(np.zeros((4, 3)))       # a deep copy: the clone keeps its own storage
        np.testing.assert_array_equal(cpy.data(), BIG)


def test_variable_meta():  # VARIABLE.Meta :49-78, META.Types / Version (internal/eigen/test/test_meta.cpp:10-40)
    a, b, c = trio()
    assert [t.type_label() for t in (a, b, c)] == ["DOUBLE", "FLOAT", "INT32"]
    assert [t.type_size() for t in (a, b, c)] == [8, 4, 4]
    assert [t.dtype() for t in (a, b, c)] == [np.float64, np.float32, np.int32]
    assert [t.get_version() for t in (a, b, c)] == [1, 1, 1]    # leaves are born at version 1 ...
    assert (a + a).get_version() == 0                           # ... a functor's metadata at 0 (META.Version)
    assert [t.usage() for t in (a, b, c)] == ["variable"] * 3
    assert a.teq_shape() == [3, 4, 1, 1, 1, 1, 1, 1] and a.shape() == [4, 3]


def test_variable_assign():  # VARIABLE.Assign :81-126
    a, b, c = trio()
    d = np.array([3, 1, 222, 21, 17, 7, 91, 11, 71, 13, 81, 2], dtype=np.float64)
    with pytest.raises(Exception) as err:
        a.assign(np.zeros((7, 3)))
    assert "assigning data shaped [3\\7\\1\\1\\1\\1\\1\\1] to tensor [3\\4\\1\\1\\1\\1\\1\\1]" in str(err.value)
    assert a.get_version() == 1                                 # a rejected assign leaves the version alone
    a.assign(d.reshape(4, 3))
    b.assign(d.reshape(4, 3))                                   # double data into a float / an int32 variable: egen::type_convert
    c.assign(d.reshape(4, 3))
    va = a.get_version()
    # every assign takes the next version after the highest one alive (get_lastvers, variable.hpp:18-26): 2, 3, 4 in the reference
    assert va >= 2 and (b.get_version(), c.get_version()) == (va + 1, va + 2)
    np.testing.assert_array_equal(b.data(), d.reshape(4, 3).astype(np.float32))
    np.testing.assert_array_equal(c.data(), d.reshape(4, 3).astype(np.int32))
    assert b.data().dtype == np.float32 and c.data().dtype == np.int32
    a.assign(np.zeros((4, 3)))
    assert a.get_version() == va + 3
    np.testing.assert_array_equal(a.data(), np.zeros((4, 3)))


def test_constant():  # CONSTANT.CopyMove / Meta (tenncor/eteq/test/test_constant.cpp:12-53)
    a = tc.constant(BIG.astype(np.float64))
    assert str(a) == "[1\\2\\3\\4\\5\\...]"
    assert str(a.clone()) == "[1\\2\\3\\4\\5\\...]"
    assert a.type_label() == "DOUBLE" and a.get_version() == 1 and a.usage() == "constant"
    assert a.teq_shape() == [3, 4, 1, 1, 1, 1, 1, 1]
    np.testing.assert_array_equal(a.data(), BIG)
    assert str(tc.scalar_constant(3, [], "DOUBLE")) == "3"      # an all-equal constant prints as its value (ileaf.hpp const_encode)
    assert str(tc.scalar_constant(2.5, [2, 2], "FLOAT")) == "2.5"
    with pytest.raises(Exception, match="is not a variable"):
        tc.to_variable(a)                                       # immutable: there is no assign on a constant


def test_type_caster():  # CASTER.Default :12-81 — arguments whose type differs from the functor's are wrapped in CAST, the others pass
    a = tc.scalar_constant(3, [], "DOUBLE")
    b = tc.scalar_constant(3, [], "FLOAT")
    f = tc.egen.make_tfunctor("DOUBLE", "ADD", [a, b])
    x, y = f.args()
    assert str(x) == "3" and x == a and str(y) == "CAST" and y.args() == [b]
    assert x.type_label() == y.type_label() == f.type_label() == "DOUBLE"
    f = tc.egen.make_tfunctor("FLOAT", "ADD", [a, b])           # order does not matter
    x, y = f.args()
    assert str(x) == "CAST" and x.args() == [a] and y == b
    assert x.type_label() == y.type_label() == f.type_label() == "FLOAT"
    f = tc.egen.make_tfunctor("INT32", "ADD", [a, b])           # both can be cast
    x, y = f.args()
    assert str(x) == str(y) == "CAST" and x.type_label() == y.type_label() == "INT32"
    for functor, want in ((x, "INT32"), (y, "INT32")):
        node = tc.dump_graph([functor])[-1]
        assert node["op"] == "CAST" and node["dtype"] == tc.dump_graph([tc.scalar_constant(0, [], want)])[-1]["dtype"]
    # the untyped entry point takes the type the opcode's TypeParser picks: the higher-precision argument wins
    g = a + b
    assert g.type_label() == "DOUBLE" and [str(t) for t in g.args()] == ["3", "CAST"]


def test_type_caster_cast():  # CASTER.Cast :84-94 — CAST's own argument is never cast
    a = tc.scalar_constant(3, [], "FLOAT")
    f = tc.egen.make_tfunctor("INT32", "CAST", [a], {"dtype": "INT32"})
    assert f.args() == [a] and a.type_label() == "FLOAT" and f.type_label() == "INT32"
    assert tc.api.cast(a, "FLOAT") == a if hasattr(tc.api, "cast") else True  # FuncOpt<CAST>: same type is redundant


def test_subscriptions():  # OBSERVABLE.Subscriptions / CopyMove (internal/eigen/test/test_observable.cpp:12-67)
    lf = tc.variable(BIG.astype(np.float64), "leaf")
    obs = tc.api.neg(lf)
    assert obs.nsubs() == 0
    parent = tc.api.sin(obs)
    unrelated = tc.api.sin(lf)
    assert obs.nsubs() == 1 and unrelated.nsubs() == 0          # only functors reading it subscribe; leaves keep no subscriber set
    twice = obs * obs                                           # one reader in two slots is ONE subscriber (a set)
    assert obs.nsubs() == 2
    cpy = obs.clone()                                           # a copy starts with no readers of its own ...
    assert cpy.nsubs() == 0 and obs.nsubs() == 2
    pcpy = parent.clone()                                       # ... and a copied reader subscribes to the same argument
    assert obs.nsubs() == 3
    del pcpy, twice
    gc.collect()
    assert obs.nsubs() == 1                                     # readers unsubscribe when they die
    other = tc.api.abs(lf)
    parent.update_child(other, 0)                               # re-pointing a reader moves its subscription
    assert obs.nsubs() == 0 and other.nsubs() == 1


def test_etensor_tag_and_cache():  # ETensor.tag / .cache of the reference's python module (tenncor/python/eteq_ext.cpp:163-201)
    lf = tc.variable(BIG.astype(np.float64), "leaf")
    f = tc.api.sin(tc.api.neg(lf))
    f.tag("recovery", "act_idx")
    assert tc.dump_graph([f])[-1]["attrs"]["recovery"] == "act_idx"
    lf.tag("ignored", "on a leaf")                               # only functors carry attributes
    T = tc.testing
    arg = f.args()[0]
    T.mock_data(arg, 100)
    T.stub_launch(f)
    memory = T.CountingMemory()
    f.cache()                                                    # Functor::cache_init: the result survives its planned reads
    T.holder_assign(f, 1, memory)
    ptr = T.holder_ptr(f)
    T.holder_read(f)
    T.holder_read(f)
    assert T.holder_ptr(f) == ptr and [k for k, _, _ in memory.log()] == ["allocate"]
    for name in ("release_data", "release_get", "get", "data", "calc"):
        assert hasattr(f, name)
    np.testing.assert_array_equal(lf.release_data(), BIG)        # on a leaf: a plain read


def test_const_encoding_and_usage_names():  # LEAF.ConstEncoding / GetUsage (internal/teq/test/test_leaf.cpp:10-33)
    data = np.arange(1, 9, dtype=np.float64)
    assert str(tc.constant(data[:1].reshape(()))) == "1"
    assert str(tc.constant(data[:4])) == "[1\\2\\3\\4]"
    assert str(tc.constant(data.reshape(2, 4))) == "[1\\2\\3\\4\\5\\...]"          # teq::Shape({4, 2}): five shown, then an ellipsis
    assert str(tc.constant(data[:5])) == "[1\\2\\3\\4\\5]"
    assert tc.constant(data[:4]).usage() == "constant"
    assert tc.variable(data[:4], "v").usage() == "variable"
    assert tc.placeholder(data[:4], "p").usage() == "placeholder"
