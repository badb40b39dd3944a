"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's Eigen back end (mingkaic/tenncor,
internal/eigen/operator.hpp) used as the checker for the CUDA path. Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import this module; the product (`tenncor_b200/`) never does and fails loudly when
its CUDA library is missing.

Why a restatement and not the reference itself: the reference's numeric kernels are
Eigen 3.3.7 expression templates (third_party/repos/eigen.bzl:4-9, conanfile.py:19);
Eigen, Boost, cppkg, gRPC and protobuf-C++ are not vendored under /root/reference and are
absent from this image, so the reference cannot be compiled here.
Parity is pinned instead by the reference's own golden vectors
(internal/eigen/test/test_operator.cpp, tenncor/test/test_equation.cpp,
tenncor/test/test_api.cpp) — tests/test_oracle_golden.py checks this module against
every one of them that was transcribed into tests/golden/.

Conventions follow the reference: rank-8 column-major tensors, teq dim 0 fastest
(internal/eigen/convert.hpp:33, internal/teq/shape.hpp:56-59). A tensor is a flat 1-D
numpy array plus an 8-long shape; `nd()` views it as a C-ordered array of the REVERSED
shape, so teq rank r is numpy axis 7 - r (the same reversal the reference's pybind layer
applies, tenncor/pyutils/src/convert.cpp:20-37).
"""
import numpy as np

RANK_CAP = 8  # internal/teq/shape.hpp:45

# checker mode (default): fp32 contractions accumulate in double, so the oracle is the more
# accurate side of every comparison. baseline mode: contractions stay in the tensor's own
# dtype (sgemm), which is what the reference's Eigen path executes — used only when the
# oracle is TIMED as the CPU baseline.
_BASELINE_MODE = False


def set_baseline_mode(on):
    global _BASELINE_MODE
    _BASELINE_MODE = bool(on)


def _acc_dtype(dt):
    if dt == np.float32 and not _BASELINE_MODE:
        return np.dtype(np.float64)
    return dt

# egen::_GENERATED_OPCODE, cfg/ops.yml:63-711 in order (BAD_OP = 0)
OPCODES = [
    "BAD_OP", "IDENTITY", "ABS", "NEG", "SIN", "COS", "TAN", "EXP", "LOG", "SQRT", "ROUND",
    "SIGMOID", "TANH", "SQUARE", "CUBE", "RAND_UNIF", "REVERSE", "REDUCE_SUM", "REDUCE_PROD",
    "REDUCE_MIN", "REDUCE_MAX", "ARGMAX", "PERMUTE", "EXTEND", "RESHAPE", "SLICE", "PAD",
    "STRIDE", "SCATTER", "POW", "ADD", "SUB", "MUL", "DIV", "MIN", "MAX", "EQ", "NEQ", "LT",
    "GT", "MATMUL", "CONTRACT", "CONV", "SELECT", "CONCAT", "ASSIGN", "ASSIGN_ADD",
    "ASSIGN_SUB", "ASSIGN_MUL", "ASSIGN_DIV", "CAST",
]
OP = {name: i for i, name in enumerate(OPCODES)}

# egen::_GENERATED_DTYPE, cfg/fulltype.yml:4-34 in order (BAD_TYPE = 0)
DTYPES = {1: np.float64, 2: np.float32, 3: np.int8, 4: np.uint8, 5: np.int16, 6: np.uint16,
          7: np.int32, 8: np.uint32, 9: np.int64, 10: np.uint64}
DTYPE_CODE = {np.dtype(v): k for k, v in DTYPES.items()}


def full_shape(shape):
    s = [int(d) for d in shape][:RANK_CAP]
    return s + [1] * (RANK_CAP - len(s))


def n_elems(shape):
    n = 1
    for d in full_shape(shape):
        n *= d
    return n


def nd(flat, shape):
    """C-ordered view with reversed shape: teq rank r == numpy axis 7 - r."""
    return np.asarray(flat).reshape(full_shape(shape)[::-1])


def flat(arr):
    return np.ascontiguousarray(arr).reshape(-1)


def ax(rank):
    return RANK_CAP - 1 - rank


# ------------------------------------------------------------------ elementwise
# internal/eigen/operator.hpp:377-661 (unary), :665-987 (binary), :1050-1067 (select)
def _sigmoid(x):
    # Eigen scalar_sigmoid_op: 1 / (1 + exp(-x))  (operator.hpp:594)
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


def _int_via_double(fn):
    def wrapped(x):
        if np.issubdtype(x.dtype, np.integer):
            return fn(x.astype(np.float64)).astype(x.dtype)
        return fn(x)
    return wrapped


def _round_half_away(x):
    # Eigen numext::round == std::round (half away from zero); np.round is half-to-even
    if np.issubdtype(x.dtype, np.integer):
        return x.copy()
    return (np.sign(x) * np.floor(np.abs(x) + x.dtype.type(0.5))).astype(x.dtype)


UNARY = {
    "ABS": np.abs,
    "NEG": np.negative,
    "SIN": _int_via_double(np.sin),
    "COS": _int_via_double(np.cos),
    "TAN": _int_via_double(np.tan),
    "EXP": _int_via_double(np.exp),
    "LOG": _int_via_double(np.log),
    "SQRT": _int_via_double(np.sqrt),
    "ROUND": _round_half_away,
    "SIGMOID": _int_via_double(_sigmoid),
    "TANH": _int_via_double(np.tanh),
    "SQUARE": lambda x: x * x,
    "CUBE": lambda x: x * x * x,
}


def _pow(a, b):
    # std::pow per element (operator.hpp:673-690); integers go through double
    if np.issubdtype(a.dtype, np.integer):
        return np.power(a.astype(np.float64), b.astype(np.float64)).astype(a.dtype)
    return np.power(a, b)


def _div(a, b):
    if np.issubdtype(a.dtype, np.integer):
        out = np.zeros_like(a)
        nz = b != 0
        # C++ integer division truncates toward zero
        out[nz] = (np.trunc(a[nz].astype(np.float64) / b[nz].astype(np.float64))).astype(a.dtype)
        return out
    return a / b


BINARY = {
    "POW": _pow,
    "ADD": lambda a, b: a + b,
    "SUB": lambda a, b: a - b,
    "MUL": lambda a, b: a * b,
    "DIV": _div,
    "MIN": np.minimum,
    "MAX": np.maximum,
    "EQ": lambda a, b: (a == b).astype(a.dtype),
    "NEQ": lambda a, b: (a != b).astype(a.dtype),
    "LT": lambda a, b: (a < b).astype(a.dtype),
    "GT": lambda a, b: (a > b).astype(a.dtype),
}


def unary(op, x):
    with np.errstate(all="ignore"):
        return UNARY[op](np.asarray(x))


def binary(op, a, b):
    with np.errstate(all="ignore"):
        return BINARY[op](np.asarray(a), np.asarray(b))


def nnary(op, args):
    # operator.hpp:716-731 / :786-794: out = args[0]; out op= args[i]
    out = np.array(args[0], copy=True)
    for a in args[1:]:
        out = out + a if op == "ADD" else out * a
    return out


def select(cond, then, otherwise):
    return np.where(np.asarray(cond) != 0, then, otherwise)


def cast(x, dtype_code):
    # operator.hpp:1239-1260, Eigen cast<T>() == static_cast
    with np.errstate(all="ignore"):
        return np.asarray(x).astype(DTYPES[dtype_code])


def rand_unif(lo, hi, rng):
    """operator.hpp:993-1044. The reference draws from std::default_random_engine, whose
    stream is implementation defined; parity is statistical (range + moments) only."""
    lo = np.asarray(lo)
    hi = np.asarray(hi)
    if np.issubdtype(lo.dtype, np.integer):
        return rng.integers(lo, hi, endpoint=True).astype(lo.dtype)
    return (lo + rng.random(lo.shape) * (hi - lo)).astype(lo.dtype)


# ------------------------------------------------------------------ reductions
# operator.hpp:54-132; out shape keeps rank with 1s (cfg/ops.yml:123-139)
_RED = {"REDUCE_SUM": np.sum, "REDUCE_PROD": np.prod, "REDUCE_MIN": np.min, "REDUCE_MAX": np.max}


def reduce(op, x, shape, ranks):
    a = nd(x, shape)
    axes = tuple(sorted(ax(r) for r in set(ranks) if r < RANK_CAP))
    if not axes:
        return flat(a).copy(), full_shape(shape)
    kw = {"dtype": a.dtype} if op in ("REDUCE_SUM", "REDUCE_PROD") else {}
    out = _RED[op](a, axis=axes, keepdims=True, **kw)
    oshape = full_shape(shape)
    for r in ranks:
        oshape[r] = 1
    return flat(out), oshape


def argmax(x, shape, return_dim):
    """operator.hpp:136-155. return_dim >= 8: flat column-major index of the max;
    otherwise the index along return_dim. First (lowest) index wins ties; result is cast
    to the tensor's own dtype."""
    a = nd(x, shape)
    if return_dim >= RANK_CAP:
        return np.array([np.argmax(flat(a))], dtype=a.dtype), full_shape([])
    out = np.argmax(a, axis=ax(return_dim))
    oshape = full_shape(shape)
    oshape[return_dim] = 1
    return flat(out.astype(a.dtype)), oshape


# ------------------------------------------------------------------ layout ops
def extend(x, shape, bcast):
    # operator.hpp:159-173: Eigen broadcast(coord) tiles each rank bcast[r] times
    b = full_shape(bcast)
    a = nd(x, shape)
    out = np.tile(a, b[::-1])
    return flat(out), [s * m for s, m in zip(full_shape(shape), b)]


def extend_bcast_from_like(in_shape, like_shape):
    # eigen::unpack_extend with a "tensor" attr (internal/eigen/src/packattr.cpp:38-53)
    i, t = full_shape(in_shape), full_shape(like_shape)
    return [t[r] if i[r] != t[r] else 1 for r in range(RANK_CAP)]


def complete_order(order):
    # missing ranks are appended in order (operator.hpp:183-197)
    order = [int(o) for o in order][:RANK_CAP]
    seen = set(order)
    return order + [r for r in range(RANK_CAP) if r not in seen]


def permute(x, shape, order):
    # out rank i = in rank order[i] (cfg/ops.yml:232-238, Eigen shuffle)
    order = complete_order(order)
    a = nd(x, shape)
    axes = [ax(order[RANK_CAP - 1 - k]) for k in range(RANK_CAP)]
    out = np.transpose(a, axes)
    s = full_shape(shape)
    return flat(out), [s[order[r]] for r in range(RANK_CAP)]


def slice_(x, shape, extents):
    """operator.hpp:214-255 with the offset/extent clamping of :225-229."""
    s = full_shape(shape)
    a = nd(x, shape)
    idx = [slice(None)] * RANK_CAP
    oshape = list(s)
    for r, (off, ext) in enumerate(list(extents)[:RANK_CAP]):
        off = min(int(off), s[r] - 1)
        ext = min(int(ext), s[r] - off)
        idx[ax(r)] = slice(off, off + ext)
        oshape[r] = ext
    return flat(a[tuple(idx)]), oshape


def pad(x, shape, paddings):
    # operator.hpp:259-275 zero fill
    a = nd(x, shape)
    pw = [(0, 0)] * RANK_CAP
    oshape = full_shape(shape)
    for r, (lo, hi) in enumerate(list(paddings)[:RANK_CAP]):
        pw[ax(r)] = (int(lo), int(hi))
        oshape[r] += int(lo) + int(hi)
    return flat(np.pad(a, pw)), oshape


def stride(x, shape, incrs):
    # operator.hpp:279-293
    a = nd(x, shape)
    idx = [slice(None)] * RANK_CAP
    for r, inc in enumerate(list(incrs)[:RANK_CAP]):
        idx[ax(r)] = slice(None, None, int(inc))
    out = a[tuple(idx)]
    return flat(out), list(out.shape[::-1])


def scatter(x, shape, out_shape, incrs):
    # operator.hpp:298-314: out.setZero(); out.stride(incrs) = in
    a = nd(x, shape)
    out = np.zeros(full_shape(out_shape)[::-1], dtype=a.dtype)
    idx = [slice(None)] * RANK_CAP
    for r, inc in enumerate(list(incrs)[:RANK_CAP]):
        idx[ax(r)] = slice(None, None, int(inc))
    out[tuple(idx)] = a
    return flat(out), full_shape(out_shape)


def reverse(x, shape, ranks):
    a = nd(x, shape)
    axes = tuple(ax(r) for r in set(ranks))
    return flat(np.flip(a, axis=axes) if axes else a), full_shape(shape)


def concat(xs, shapes, axis):
    # operator.hpp:336-368: binary concatenate, or n-ary chips of extent 1
    arrs = [nd(x, s) for x, s in zip(xs, shapes)]
    out = np.concatenate(arrs, axis=ax(axis))
    return flat(out), list(out.shape[::-1])


# ------------------------------------------------------------------ contractions
def contract(a, ashape, b, bshape, pairs):
    """operator.hpp:1069-1101: out ranks = b-free (in order) then a-free
    (cfg/ops.yml:547-565); `pairs` are (a_rank, b_rank)."""
    A, B = nd(a, ashape), nd(b, bshape)
    pairs = [(int(p), int(q)) for p, q in pairs if p < RANK_CAP and q < RANK_CAP]
    a_axes = [ax(p) for p, _ in pairs]
    b_axes = [ax(q) for _, q in pairs]
    acc = _acc_dtype(A.dtype)
    out = np.tensordot(A.astype(acc), B.astype(acc), axes=(a_axes, b_axes)).astype(A.dtype)
    # tensordot(A, B) axes: A-free (numpy order = teq ranks descending) then B-free (descending)
    # reading numpy axes right-to-left gives teq order: b-free ascending, then a-free ascending
    as_, bs_ = full_shape(ashape), full_shape(bshape)
    acom = {p for p, _ in pairs}
    bcom = {q for _, q in pairs}
    oshape = [bs_[r] for r in range(RANK_CAP) if r not in bcom and bs_[r] != 1]
    oshape += [as_[r] for r in range(RANK_CAP) if r not in acom and as_[r] != 1]
    assert n_elems(oshape) == out.size
    return flat(out), full_shape(oshape)


def matmul(a, ashape, b, bshape):
    """operator.hpp:1108-1139: C[N,M,...] = A[K,M,...] . B[N,K,...]; batched over ranks 2+."""
    as_, bs_ = full_shape(ashape), full_shape(bshape)
    A = nd(a, ashape).reshape(-1, as_[1], as_[0])  # row-major (M x K) per batch
    B = nd(b, bshape).reshape(-1, bs_[1], bs_[0])  # row-major (K x N)
    acc = _acc_dtype(A.dtype)
    out = np.matmul(A.astype(acc), B.astype(acc)).astype(A.dtype)
    return flat(out), full_shape([bs_[0], as_[1]] + as_[2:])


def conv(img, ishape, kern, kshape, order):
    """operator.hpp:1143-1187: N-d VALID CORRELATION (no flip, no stride): kernel rank i
    slides along image rank order[i] (golden test_operator.cpp:2630-2634)."""
    order = complete_order(order)
    I, K = nd(img, ishape), nd(kern, kshape)
    is_, ks_ = full_shape(ishape), full_shape(kshape)
    oshape = list(is_)
    for i in range(RANK_CAP):
        oshape[order[i]] = is_[order[i]] - ks_[i] + 1
    acc = _acc_dtype(I.dtype)
    out = np.zeros(oshape[::-1], dtype=acc)
    for kidx in np.ndindex(*ks_):  # kidx[i] = coordinate along kernel rank i
        sl = [None] * RANK_CAP
        for i in range(RANK_CAP):
            r = order[i]
            sl[ax(r)] = slice(kidx[i], kidx[i] + oshape[r])
        kval = K[tuple(kidx[::-1])]
        out += I[tuple(sl)].astype(acc) * np.asarray(kval).astype(acc)
    return flat(out.astype(I.dtype)), oshape


# ------------------------------------------------------------------ assign
def assign(op, target, source):
    # eigen::TensAssign, device.hpp:507-544; operator.hpp:1190-1237 (in place)
    if op == "ASSIGN":
        target[...] = source
    elif op == "ASSIGN_ADD":
        target += source
    elif op == "ASSIGN_SUB":
        target -= source
    elif op == "ASSIGN_MUL":
        target *= source
    elif op == "ASSIGN_DIV":
        target[...] = _div(target, source)
    else:
        raise ValueError(op)
    return target


# ------------------------------------------------------------------ graph (tape) evaluation
def eval_tape(tape, rng=None):
    """Evaluate a dumped functor graph node by node, the way the reference does:
    post-order, one unfused op per functor (teq::TravEvaluator::visit_func,
    internal/teq/evaluator.hpp:34-43 -> eigen::Device::calc, internal/eigen/device.hpp:555-570),
    ASSIGN* writing in place into the variable's storage.

    `tape` is a list of node dicts in evaluation order (as produced by the product's
    `tenncor_b200.dump_graph`):
      leaf:    {"id", "kind": "leaf", "shape", "dtype", "data": flat ndarray (shared, mutable)}
      functor: {"id", "kind": "func", "op": name, "args": [ids], "shape", "dtype", "attrs": {...}}
    Returns {id: flat ndarray}.
    """
    rng = rng or np.random.default_rng(0)
    val = {}
    shp = {}
    for node in tape:
        nid = node["id"]
        shape = full_shape(node["shape"])
        shp[nid] = shape
        if node["kind"] == "leaf":
            val[nid] = node["data"]
            continue
        op = node["op"]
        args = node["args"]
        at = node.get("attrs", {})
        a = [val[i] for i in args]
        s = [shp[i] for i in args]
        if op in ("IDENTITY", "RESHAPE"):
            out = a[0]  # alias (eigen::ref, src/operator.cpp:12-15)
        elif op in UNARY:
            out = unary(op, a[0])
        elif op in ("ADD", "MUL"):
            out = nnary(op, a)
        elif op in BINARY:
            out = binary(op, a[0], a[1])
        elif op == "SELECT":
            out = select(a[0], a[1], a[2])
        elif op == "CAST":
            out = cast(a[0], node["dtype"])
        elif op == "RAND_UNIF":
            out = rand_unif(a[0], a[1], rng)
        elif op in _RED:
            out, _ = reduce(op, a[0], s[0], at["rank_set"])
        elif op == "ARGMAX":
            out, _ = argmax(a[0], s[0], at["rank"])
        elif op == "EXTEND":
            bc = at["dimensions"] if "dimensions" in at else extend_bcast_from_like(s[0], at["tensor_shape"])
            out, _ = extend(a[0], s[0], bc)
        elif op == "PERMUTE":
            out, _ = permute(a[0], s[0], at["ranks"])
        elif op == "SLICE":
            out, _ = slice_(a[0], s[0], at["dimension_pairs"])
        elif op == "PAD":
            out, _ = pad(a[0], s[0], at["dimension_pairs"])
        elif op == "STRIDE":
            out, _ = stride(a[0], s[0], at["dimensions"])
        elif op == "SCATTER":
            out, _ = scatter(a[0], s[0], at["shape"], at["dimensions"])
        elif op == "REVERSE":
            out, _ = reverse(a[0], s[0], at["rank_set"])
        elif op == "CONCAT":
            out, _ = concat(a, s, at["rank"])
        elif op == "MATMUL":
            out, _ = matmul(a[0], s[0], a[1], s[1])
        elif op == "CONTRACT":
            out, _ = contract(a[0], s[0], a[1], s[1], at["rank_pairs"])
        elif op == "CONV":
            out, _ = conv(a[0], s[0], a[1], s[1], at["ranks"])
        elif op.startswith("ASSIGN"):
            out = assign(op, a[0], a[1])
        else:
            raise ValueError("oracle: unknown op %s" % op)
        out = np.asarray(out)
        if out.dtype != DTYPES[node["dtype"]] and not op.startswith("ASSIGN") and op not in ("IDENTITY", "RESHAPE"):
            out = out.astype(DTYPES[node["dtype"]])
        assert out.size == n_elems(shape), (op, out.size, shape)
        val[nid] = out
    return val
