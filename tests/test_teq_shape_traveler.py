"""teq::Shape and the graph travelers under the evaluators and teq::derive, mirrored from
internal/teq/test/test_shape.cpp and test_traveler.cpp. Real functors stand in for the reference's mock functors; host only."""
import numpy as np
import pytest

import tenncor_b200 as tc


@pytest.fixture(autouse=True)
def _built(built):
    tc.require_host()


Shape = lambda *a: tc.teq.Shape(*a)  # noqa: E731
CAP = 8


def scalar(name):
    return tc.variable(np.ones((), dtype=np.float64), name)


# ------------------------------------------------------------------------------------------------ test_shape.cpp

def test_shape_init():  # SHAPE.Init :18-58
    assert tc.teq.rank_cap == CAP
    sc = Shape()
    assert [sc.at(i) for i in range(CAP)] == [1] * CAP
    vec = Shape([12, 43, 56])
    assert [vec.at(i) for i in range(CAP)] == [12, 43, 56, 1, 1, 1, 1, 1]
    longlist = [4, 23, 44, 52, 19, 92, 12, 2, 5]
    assert Shape(longlist).to_list() == longlist[:CAP]           # dimensions beyond rank_cap are dropped
    with pytest.raises(Exception) as err:
        Shape([43, 2, 5, 33, 0, 2, 7])
    assert "cannot create shape with vector containing zero: [43\\2\\5\\33\\0\\2\\7]" in str(err.value)
    for s in (sc, vec):
        with pytest.raises(Exception, match="cannot access out of bounds index 8"):
            s.at(CAP)


def test_shape_assign_and_copy():  # SHAPE.VecAssign / Moves / Iterators :61-129
    slist = [52, 58, 35, 46, 77, 80]
    s = Shape(slist)
    assert len(s) == CAP and s.to_list() == slist + [1, 1]
    assert Shape(s.to_list()) == s and not (Shape([7, 42]) == s)
    with pytest.raises(Exception) as err:
        Shape([3, 0, 11, 89])
    assert "cannot create shape with vector containing zero: [3\\0\\11\\89]" in str(err.value)


def test_shape_nelems():  # SHAPE.NElems :132-147 — 255^8 still fits the element counter
    assert Shape([11, 12, 16]).n_elems() == 11 * 12 * 16
    assert Shape([255] * 8).n_elems() == 17878103347812890625
    assert Shape([65536, 784]).n_elems() == 65536 * 784          # DimT is 32 bits here (the reference's -DSDIM_BYTES=4): batch 65536 fits one rank


def test_shape_compatible():  # SHAPE.Compatible :150-194
    slist = [20, 48, 10, 27, 65, 74]
    shape = Shape(slist)
    assert all(shape.compatible_after(shape, idx) for idx in range(CAP))
    at = 3
    ilist = slist[:at] + [2] + slist[at:]
    ishape = Shape(ilist)
    assert not any(shape.compatible_after(ishape, idx) for idx in range(at))
    ilist[at] = 3
    ishape2 = Shape(ilist)
    assert not any(ishape.compatible_after(ishape2, idx) for idx in range(at + 1))
    assert all(ishape.compatible_after(ishape2, idx) for idx in range(at + 1, CAP))
    assert all(ishape.compatible_before(ishape2, idx) for idx in range(at + 1))
    assert not any(ishape.compatible_before(ishape2, idx) for idx in range(at + 1, CAP + 1))


def test_shape_to_string_and_narrow():  # SHAPE.ToString / NarrowShape :197-220
    assert str(Shape([24, 11, 12, 16, 7, 71, 1, 1])) == "[24\\11\\12\\16\\7\\71\\1\\1]"
    assert Shape([1, 2, 3, 4, 1]).narrow() == [1, 2, 3, 4]
    assert Shape().narrow() == []


# ------------------------------------------------------------------------------------------------ test_traveler.cpp

def test_graph_stat():  # TRAVELER.GraphStat :16-33 — height = longest distance to a leaf
    a, b, c = scalar("a"), scalar("b"), scalar("c")
    d = tc.api.neg(c)
    f = a + b
    g = d * f
    g2 = tc.api.sin(g) - d                                       # longest path wins: sin(g) is 3 above the leaves, d only 1
    height = dict(tc.teq.graph_stat(g2))
    assert [height[t] for t in (a, b, c, d, f, g, g2)] == [0, 0, 0, 1, 1, 2, 4]
    assert len(height) == 8


def test_graph_index():  # traveler.hpp:112-148 — post-order numbering, the tie-break of teq::derive's accumulation order
    a, b, c = scalar("a"), scalar("b"), scalar("c")
    d = tc.api.neg(c)
    f = a + b
    g = d * f
    index = dict(tc.teq.graph_index(g))
    assert [index[t] for t in (c, d, a, b, f, g)] == [0, 1, 2, 3, 4, 5]


def roadmap(root, targets, follow_attrs=True):
    return {t: (args, attrs) for t, args, attrs in tc.teq.path_finder(root, targets, follow_attrs)}


def test_path_finder():  # TRAVELER.PathFinder :36-107
    a, b, c = scalar("a"), scalar("b"), scalar("c")
    d = tc.api.neg(c)
    f = a + b
    g = d * f
    for_a, for_d = roadmap(g, [a]), roadmap(g, [d])
    assert for_a[g] == ([1], []) and for_d[g] == ([0], [])       # through which arguments a target is reached
    assert for_a[f] == ([0], [])
    assert d not in for_a and d not in for_d                     # a target itself, and branches without one, stay off the map
    both = roadmap(g, [a, d])
    assert both[g] == ([0, 1], []) and both[f] == ([0], [])
    sub = roadmap(f, [a])                                        # started lower: nothing above the start is known
    assert g not in sub and sub[f] == ([0], [])
    assert roadmap(g, [c]) == {g: ([0], []), d: ([0], [])}
    assert roadmap(f, [c]) == {}


def test_path_finder_attr():  # TRAVELER.PathFinderAttr :110-155 — tensors referenced by attributes are followed too
    a, b, c = scalar("a"), scalar("b"), scalar("c")
    d = tc.api.extend_like(b, c)                                 # EXTEND keeps `c` as its `tensor` attribute
    f = a + c
    g = tc.egen.make_functor("ADD", [a, d], {"tensor": f})
    assert tc.teq.attr_tensors(d) == [c] and tc.teq.attr_tensors(g) == [f] and tc.teq.attr_tensors(f) == []
    rm = roadmap(g, [c])
    assert rm[d] == ([], ["tensor"])                             # no argument leads to c, the attribute does
    assert rm[g] == ([1], ["tensor"])                            # argument d (through its attribute) and the attribute f
    assert rm[f] == ([1], [])
    assert roadmap(g, [c], follow_attrs=False).keys() == {g, d}  # attribute tensors are still matched, only not walked into


def test_copier():  # Copier, traveler.hpp:378-432 (what layr::deep_clone stands on): ignored nodes are shared, the rest cloned once
    a, b, c = scalar("a"), scalar("b"), scalar("c")
    f = a + b
    g = f * c
    h = g - f                                                    # f is reached twice
    clones = dict(tc.teq.copy_graph(h, [c]))
    assert set(clones) == {a, b, f, g, h}
    assert all(orig != cpy for orig, cpy in clones.items())
    hc = clones[h]
    assert [str(t) for t in hc.args()] == ["MUL", "ADD"]
    assert hc.args()[0].args() == [clones[f], c]                 # the ignored leaf is the original object
    assert hc.args()[1] == clones[f]                             # one clone per node, shared like the original
    assert clones[f].args() == [clones[a], clones[b]]
    assert h.args() == [g, f] and f.nsubs() == 2                 # the source graph is untouched
