"""The holder / device contract of SURVEY.md §8 rows a2-a7 and a16, mirrored from internal/eigen/test/test_device.cpp and
tenncor/eteq/test/test_functor.cpp.

The reference tests hand `TensOp` a lambda and `assign` a MockRuntimeMemory and check WHO is called with WHICH pointers and
lifetimes. Here a recording launcher (tc.testing.stub_launch) replaces the kernel launch and a recording allocator
(tc.testing.CountingMemory) the arena, so the whole contract runs without a GPU: when a holder allocates, what it hands the
kernel, how many consumer reads a result survives (Expirable, internal/eigen/memory.hpp:57-139), and when
eigen::Device::calc (internal/eigen/device.hpp:555-570) recomputes, extends or leaves a node alone."""
import numpy as np
import pytest

import tenncor_b200 as tc

T = tc.testing


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


def leaf(shape=(3, 4), name="leaf", dtype=np.float64):
    return tc.variable(np.arange(int(np.prod(shape)), dtype=dtype).reshape(shape), name)


def resident_arg(ttl=100, shape=(3, 4), dtype=np.float64):
    """a functor argument that already holds (host-mocked) data, like the reference's make_obs(data, devref, ...)"""
    arg = tc.api.neg(leaf(shape, dtype=dtype))
    T.mock_data(arg, ttl)
    return arg


# ------------------------------------------------------------------------------------------------ test_device.cpp

def test_src_ref():  # DEVICE.SrcRef :15-37 — leaf storage: data readable, assign() neither allocates nor changes it
    data = np.array([[1, 2], [3, 4]], dtype=np.float64)
    var = tc.variable(data, "src")
    assert T.holder_kind(var) == "DevSrc"
    np.testing.assert_array_equal(var.data(), data)
    memory = T.CountingMemory()
    T.holder_assign(var, 0, memory)  # referencing shouldn't do anything
    assert memory.log() == []
    np.testing.assert_array_equal(var.data(), data)


def test_tens_op():  # DEVICE.TensOp / MatOp :75-153
    arg = resident_arg()
    f = tc.api.sin(arg)
    assert T.holder_kind(f) == "DevOp"
    launches = T.stub_launch(f)
    assert launches.calls() == []          # EXPECT_FALSE(init_called)
    assert T.holder_ptr(f) == 0            # EXPECT_EQ(nullptr, ref.data())
    memory = T.CountingMemory()
    outbytes = 12 * 8
    T.holder_assign(f, 1, memory)          # should initialize
    (kind, size, ptr), = memory.log()
    assert (kind, size) == ("allocate", outbytes)
    (out, args), = launches.calls()
    assert out == ptr == T.holder_ptr(f)   # the kernel writes the block the allocator returned ...
    assert args == [T.holder_ptr(arg)]     # ... and reads the argument's own buffer, no copy
    del f                                  # the holder dies with its functor: the block goes back, with its size
    import gc; gc.collect()
    assert memory.log()[1] == ("deallocate", outbytes, ptr) and len(memory.log()) == 2


def test_expirable_ttl():  # memory.hpp:57-139 through the holder: a result survives exactly `ttl` consumer reads
    f = tc.api.sin(resident_arg())
    T.stub_launch(f)
    memory = T.CountingMemory()
    T.holder_assign(f, 2, memory)
    ptr = T.holder_ptr(f)
    assert T.holder_valid_for(f, 2) and not T.holder_valid_for(f, 3)
    assert T.holder_read(f) == ptr         # first consumer done: ttl 2 -> 1
    assert T.holder_ptr(f) == ptr and T.holder_valid_for(f, 1) and not T.holder_valid_for(f, 2)
    assert T.holder_read(f) == ptr         # second: ttl -> 0, expire()
    assert T.holder_ptr(f) == 0
    assert memory.log() == [("allocate", 96, ptr), ("deallocate", 96, ptr)]
    with pytest.raises(Exception, match="cannot extend ttl of expired Expirable"):
        T.holder_extend_life(f, 1)
    T.holder_assign(f, 1, memory)          # an expired holder borrows again on the next assign
    assert [k for k, _, _ in memory.log()] == ["allocate", "deallocate", "allocate"]
    T.holder_extend_life(f, 5)             # extend_life only ever raises the ttl
    T.holder_extend_life(f, 3)
    assert T.holder_valid_for(f, 5) and not T.holder_valid_for(f, 6)
    T.holder_assign(f, 2, memory)          # re-assign of a live holder re-runs the kernel in the same block
    assert [k for k, _, _ in memory.log()] == ["allocate", "deallocate", "allocate"]


def test_argument_without_data_is_fatal():
    arg = tc.api.neg(leaf())
    f = tc.api.sin(arg)
    T.stub_launch(arg)
    T.stub_launch(f)
    with pytest.raises(Exception, match="argument NEG has no data"):
        T.holder_assign(f, 1, T.CountingMemory())


def test_tens_ref_forwards_ticks():  # TensRef, device.hpp:383-436: an alias owns no memory; its last read is one read of the referent
    src = tc.api.sin(resident_arg())
    T.stub_launch(src)
    memory = T.CountingMemory()
    alias = tc.api.identity(src) if hasattr(tc.api, "identity") else tc.egen.make_functor("IDENTITY", [src])
    assert T.holder_kind(alias) == "DevRef"
    T.holder_assign(src, 1, memory)
    ptr = T.holder_ptr(src)
    T.holder_assign(alias, 2, memory)      # no allocation, no launch: only a lifetime
    assert len(memory.log()) == 1
    assert T.holder_ptr(alias) == ptr
    assert T.holder_valid_for(alias, 2) and not T.holder_valid_for(alias, 3)
    assert T.holder_read(alias) == ptr     # alias ttl 2 -> 1, the referent is untouched
    assert T.holder_ptr(src) == ptr
    assert T.holder_read(alias) == ptr     # alias ttl -> 0: forwards ONE tick, which exhausts the referent's single life
    assert T.holder_ptr(src) == 0 and T.holder_ptr(alias) == 0
    assert memory.log()[-1] == ("deallocate", 96, ptr)


def test_calc():  # DEVICE.Calc :156-208 — recompute iff the version propagates or there is no data; ttl = max(1, nsubs + is_target)
    a = resident_arg()
    var = leaf(name="v")
    obs = a + tc.api.neg(var)
    T.mock_data(obs.args()[1], 100)
    launches = T.stub_launch(obs)
    memory = T.CountingMemory()

    T.device_calc(obs, 0, memory)                        # first visit: no data -> assign(max(1, 0 + 0))
    assert len(launches.calls()) == 1
    assert T.holder_valid_for(obs, 1) and not T.holder_valid_for(obs, 2)

    T.device_calc(obs, 0, memory, max_version=0)         # nothing to propagate and data present: left alone
    assert len(launches.calls()) == 1
    T.device_calc(obs, 1, memory)                        # as a target it must survive one more read: extend, do not recompute
    assert len(launches.calls()) == 1
    assert T.holder_valid_for(obs, 1)

    var.assign(np.ones((3, 4)))                          # a leaf below it changed ...
    stale = obs.args()[1]
    assert stale.prop_version()                          # (its direct reader first, like the post-order walk does)
    T.device_calc(obs, 0, memory, max_version=10)        # ... so the version propagates: assign again
    assert len(launches.calls()) == 2
    T.device_calc(obs, 0, memory, max_version=10)
    assert len(launches.calls()) == 2

    p1, p2 = tc.api.sin(obs), tc.api.cos(obs)            # two subscribers (MockMObservable parent, parent2)
    assert obs.nsubs() == 2
    var.assign(np.zeros((3, 4)))
    assert stale.prop_version()
    T.device_calc(obs, 0, memory)                        # assign(2, _): one life per reader
    assert len(launches.calls()) == 3
    assert T.holder_valid_for(obs, 2) and not T.holder_valid_for(obs, 3)
    T.device_calc(obs, 1, memory)                        # readers + target read
    assert len(launches.calls()) == 3
    assert T.holder_valid_for(obs, 3) and not T.holder_valid_for(obs, 4)
    del p1, p2


def test_calc_recomputes_expired_results():
    """the 'stateless device' clause (device.hpp:561-563): same version but the buffer was consumed -> assign again"""
    obs = tc.api.sin(resident_arg())
    launches = T.stub_launch(obs)
    memory = T.CountingMemory()
    T.device_calc(obs, 1, memory)
    T.holder_read(obs)
    assert T.holder_ptr(obs) == 0
    T.device_calc(obs, 1, memory)
    assert len(launches.calls()) == 2 and T.holder_ptr(obs) != 0


# ------------------------------------------------------------------------------------------------ test_functor.cpp

def test_functor_initiation():  # FUNCTOR.Initiation :21-112
    # through make_funcattr the TypeParser meets the empty list first (eigen::no_argument_err, packattr.hpp:11)
    with pytest.raises(Exception, match="cannot operate without inputs"):
        tc.egen.make_functor("ADD", [])
    lf = leaf((3, 4))
    a, b = tc.api.neg(lf), tc.api.abs(lf)
    f = a + b
    g = tc.api.sin(f)
    assert f.teq_shape() == [4, 3, 1, 1, 1, 1, 1, 1]
    assert f.opname() == "ADD" and str(f) == "ADD"
    # holders are created eagerly here (the reference's SKIP_INIT build defers them); uninitialize drops this node's and every reader's
    assert f.has_data() and g.has_data()
    f.uninitialize()
    assert not f.has_data() and not g.has_data()
    assert a.has_data() and b.has_data()
    g.must_initialize()                                  # initializes the arguments it needs first
    assert g.has_data() and f.has_data()
    a.uninitialize()
    assert not a.has_data() and not f.has_data() and not g.has_data() and b.has_data()
    g.must_initialize()
    assert a.has_data() and f.has_data() and g.has_data()
    fcpy = f.clone()
    assert fcpy.has_data() and fcpy.args() == f.args() and fcpy != f
    assert a.nsubs() == 2                                # the copy subscribed to the same arguments


def test_functor_update_child():  # FUNCTOR.UpdateChild :115-196
    lf = leaf((3, 4))
    a, b, c = tc.api.neg(lf), tc.api.abs(lf), tc.api.sin(lf)
    f = a + b
    assert f.args() == [a, b]
    f.update_child(c, 1)
    assert not f.has_data()
    assert f.args() == [a, c]
    assert b.nsubs() == 0 and c.nsubs() == 1
    f.update_child(c, 0)
    assert not f.has_data()
    assert f.args() == [c, c]
    d = tc.api.neg(leaf((4, 3)))
    with pytest.raises(Exception) as err:
        f.update_child(d, 1)
    assert ("cannot update child 1 to argument with incompatible shape [3\\4\\1\\1\\1\\1\\1\\1] "
            "(requires shape [4\\3\\1\\1\\1\\1\\1\\1])") in str(err.value)
    e = leaf((3, 4), dtype=np.float32)
    with pytest.raises(Exception) as err:
        f.update_child(e, 0)
    assert "cannot update child 0 to argument with different type FLOAT (requires type DOUBLE)" in str(err.value)
    with pytest.raises(Exception) as err:
        f.update_child(a, 2)
    assert "cannot replace argument 2 when only there are only 2 available" in str(err.value)


def test_functor_prop_version():  # FUNCTOR.Prop :199-262
    va, vb = leaf(name="a"), leaf(name="b")
    f = va + vb
    assert f.get_version() == 0
    va.assign(np.ones((3, 4)))
    vb.assign(np.ones((3, 4)))
    newest = max(va.get_version(), vb.get_version())
    assert newest >= 1
    assert f.prop_version(newest + 2)                    # takes the newest argument version, once
    assert f.get_version() == newest
    assert not f.prop_version(newest + 2)
    # a non-idempotent opcode wants to run on every visit: its version climbs by one per call until max_version stops it
    g = tc.egen.make_functor("RAND_UNIF", [va, vb])
    cap = newest + 2
    assert g.prop_version(cap) and g.get_version() == newest
    assert g.prop_version(cap) and g.get_version() == newest + 1
    assert g.prop_version(cap) and g.get_version() == newest + 2
    assert not g.prop_version(cap)
    # max_version caps an idempotent node below its arguments' version too
    h = va * vb
    assert h.prop_version(1) and h.get_version() == 1
    assert not h.prop_version(1)
    assert h.prop_version() and h.get_version() == newest


def test_functor_cache():  # FUNCTOR.Cache :265-300 — a cached functor's result outlives its consumers' reads
    if not hasattr(T, "cache_init"):
        pytest.skip("cache_init hook not bound")
    f = tc.api.sin(resident_arg())
    T.stub_launch(f)
    T.cache_init(f)
    memory = T.CountingMemory()
    T.holder_assign(f, 1, memory)
    ptr = T.holder_ptr(f)
    T.holder_read(f)
    assert T.holder_ptr(f) == ptr                        # EXPECT_NE(nullptr, dev->data()) after the only planned read
    assert [k for k, _, _ in memory.log()] == ["allocate"]
