"""tcr_gemm_grouped (gemm_rnn.cu): grouped / K-segmented small-batch product on the tensor cores against float64.

Tolerances as for tcr_gemm (tests/test_gemm_tc_gpu.py): with S = sum_k |a||b|, TF32 |err| <= 2^-10 S, 3xTF32 |err| <= (2^-19 + K 2^-23) S
on the pre-activation; activations are 1-Lipschitz (sigmoid 1/4), so the same bound holds after them, plus 1e-6 for expf / tanhf.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def run_grouped(gpu, A_segs, B, bias, acts, b_trans, precision, cell=None, accumulate=None, cluster=None):
    """A_segs: list of [m, K_s]; B[g][s]: [K_s, n] (math layout); returns list of out_g (and (c, h) when cell)."""
    lib = gpu.lib()
    groups, segs = len(B), len(A_segs)
    m = A_segs[0].shape[0]
    n = B[0][0].shape[1]
    d = gpu.GemmGroupDesc()
    d.m, d.n, d.groups, d.segments = m, n, groups, segs
    keep = []
    for s, a in enumerate(A_segs):
        pitch_a = (a.shape[1] + 3) // 4 * 4  # TMA: row pitch a multiple of 16 bytes; the padding holds garbage on purpose
        padded = np.full((m, pitch_a), 7.0, np.float32)
        padded[:, :a.shape[1]] = a
        da = gpu.to_device(padded)
        keep.append(da)
        d.a[s] = da.ptr
        d.a_pitch[s] = pitch_a
        d.seg_k[s] = a.shape[1]
    # one buffer per group holding the segments stacked along K, like a weight matrix W_g [K x n] (or its transpose)
    pitch = None
    for g in range(groups):
        full = np.concatenate([np.asarray(B[g][s], np.float32) for s in range(segs)], axis=0)  # [K, n]
        if b_trans:
            # [n x K_s] row-major per segment: keep segments as separate buffers (different weight matrices in the backward use)
            for s in range(segs):
                bt = np.ascontiguousarray(np.asarray(B[g][s], np.float32).T)
                db = gpu.to_device(bt)
                keep.append(db)
                d.b[g][s] = db.ptr
                assert pitch in (None, bt.shape[1]), "b_trans segments must share K for one pitch"
                pitch = bt.shape[1]
        else:
            db = gpu.to_device(np.ascontiguousarray(full))
            keep.append(db)
            off = 0
            for s in range(segs):
                d.b[g][s] = db.ptr + 4 * off * n
                off += A_segs[s].shape[1]
            pitch = n
    d.b_pitch = pitch
    d.b_trans = int(b_trans)
    d.precision = precision
    outs = []
    for g in range(groups):
        if bias is not None and bias[g] is not None:
            dbias = gpu.to_device(np.asarray(bias[g], np.float32))
            keep.append(dbias)
            d.bias[g] = dbias.ptr
        d.act[g] = acts[g]
        o = gpu.to_device(accumulate[g]) if accumulate is not None else gpu.empty(m * n, np.float32)
        outs.append(o)
        d.out[g] = o.ptr
    d.out_pitch = n
    d.accumulate = 1 if accumulate is not None else 0
    cbuf = hbuf = None
    if cell is not None:
        d.cell = 1
        d.role_cand, d.role_in, d.role_forget, d.role_out = cell["roles"]
        if cell.get("c_prev") is not None:
            cp = gpu.to_device(np.asarray(cell["c_prev"], np.float32))
            keep.append(cp)
            d.c_prev = cp.ptr
        cbuf, hbuf = gpu.empty(m * n, np.float32), gpu.empty(m * n, np.float32)
        d.c_out, d.h_out, d.state_pitch = cbuf.ptr, hbuf.ptr, n
    gpu.check(lib.tcr_gemm_grouped(C.byref(d)))
    res = [gpu.to_host(o, m * n, np.float32).reshape(m, n) for o in outs]
    if cell is not None:
        return res, gpu.to_host(cbuf, m * n, np.float32).reshape(m, n), gpu.to_host(hbuf, m * n, np.float32).reshape(m, n)
    return res


def reference(A_segs, B, bias, acts, gpu):
    pre, S = [], []
    for g in range(len(B)):
        x = sum(A_segs[s].astype(np.float64) @ np.asarray(B[g][s], np.float64) for s in range(len(A_segs)))
        sabs = sum(np.abs(A_segs[s]).astype(np.float64) @ np.abs(np.asarray(B[g][s], np.float64)) for s in range(len(A_segs)))
        if bias is not None and bias[g] is not None:
            x = x + np.asarray(bias[g], np.float64)[None, :]
        pre.append(x)
        S.append(sabs)
    out = []
    for g, x in enumerate(pre):
        out.append(sigmoid(x) if acts[g] == gpu.OP["SIGMOID"] else np.tanh(x) if acts[g] == gpu.OP["TANH"] else x)
    return out, S


def bound(S, K, precision):
    return S * (2.0 ** -10 if precision == 1 else (2.0 ** -19 + K * 2.0 ** -23)) + 1e-6


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("m,n,ks", [(64, 1024, (128, 1024)), (5, 96, (12, 40)), (100, 200, (33, 70)), (64, 64, (32,)), (1, 1024, (128, 1024))],
                         ids=["c4", "tiny", "ragged", "onekb", "unbatched"])
def test_gate_products_forward(gpu, m, n, ks, precision):
    """groups = 4 gates over K-segments (x_t | h_{t-1}), bias + SIGMOID / TANH per gate: the LSTM step's forward products"""
    rng = np.random.default_rng(m * 31 + n)
    A = [rng.uniform(-1, 1, (m, k)).astype(np.float32) for k in ks]
    B = [[rng.uniform(-1, 1, (k, n)).astype(np.float32) * 0.1 for k in ks] for _ in range(4)]
    bias = [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(4)]
    acts = [gpu.OP["TANH"], gpu.OP["SIGMOID"], gpu.OP["SIGMOID"], 0]
    got = run_grouped(gpu, A, B, bias, acts, 0, precision)
    want, S = reference(A, B, bias, acts, gpu)
    for g in range(4):
        err = np.abs(got[g] - want[g])
        assert np.all(err <= bound(S[g], sum(ks), precision)), (g, float(err.max()))


@pytest.mark.parametrize("groups", [1, 2])
def test_fewer_groups(gpu, groups):
    rng = np.random.default_rng(groups)
    m, n, ks = 48, 320, (64, 96)
    A = [rng.uniform(-1, 1, (m, k)).astype(np.float32) for k in ks]
    B = [[rng.uniform(-1, 1, (k, n)).astype(np.float32) for k in ks] for _ in range(groups)]
    acts = [0] * groups
    got = run_grouped(gpu, A, B, None, acts, 0, 2)
    want, S = reference(A, B, None, acts, gpu)
    for g in range(groups):
        assert np.all(np.abs(got[g] - want[g]) <= bound(S[g], sum(ks), 2))


@pytest.mark.parametrize("precision", [1, 2], ids=["tf32", "3xtf32"])
@pytest.mark.parametrize("m,n,k", [(64, 1024, 1024), (7, 100, 52), (130, 256, 64)], ids=["c4", "tiny", "two_batch_tiles"])
def test_sum_of_products_backward(gpu, m, n, k, precision):
    """groups = 1, four K-segments, transposed weights: dh = sum_g dpre_g . W_g^T, optionally accumulated onto an existing gradient"""
    rng = np.random.default_rng(m + n + k)
    A = [rng.uniform(-1, 1, (m, k)).astype(np.float32) for _ in range(4)]
    B = [[rng.uniform(-1, 1, (k, n)).astype(np.float32) * 0.1 for _ in range(4)]]
    got = run_grouped(gpu, A, B, None, [0], 1, precision)
    want, S = reference(A, B, None, [0], gpu)
    assert np.all(np.abs(got[0] - want[0]) <= bound(S[0], 4 * k, precision)), float(np.abs(got[0] - want[0]).max())
    base = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    got = run_grouped(gpu, A, B, None, [0], 1, precision, accumulate=[base.reshape(-1).copy()])
    assert np.all(np.abs(got[0] - (want[0] + base)) <= bound(S[0], 4 * k, precision) + 1e-6)


@pytest.mark.parametrize("with_state", [True, False])
def test_lstm_cell_epilogue(gpu, with_state):
    """c_t = cand * in + c_{t-1} * forget, h_t = c_t * out (cfg/tenncor/layer.yml:758-760) from the four gate accumulators"""
    rng = np.random.default_rng(17)
    m, n, ks = 64, 1024, (128, 1024)
    A = [rng.uniform(-1, 1, (m, k)).astype(np.float32) for k in ks]
    B = [[rng.uniform(-1, 1, (k, n)).astype(np.float32) * 0.05 for k in ks] for _ in range(4)]
    bias = [rng.uniform(-0.5, 0.5, n).astype(np.float32) for _ in range(4)]
    acts = [gpu.OP["TANH"], gpu.OP["SIGMOID"], gpu.OP["SIGMOID"], gpu.OP["SIGMOID"]]  # groups: cand, forget, in, out
    roles = (0, 2, 1, 3)
    c_prev = rng.uniform(-1, 1, (m, n)).astype(np.float32) if with_state else None
    got, c, h = run_grouped(gpu, A, B, bias, acts, 0, 2, cell={"roles": roles, "c_prev": c_prev})
    want, S = reference(A, B, bias, acts, gpu)
    cw = want[0] * want[2] + (c_prev.astype(np.float64) if with_state else 0.0) * want[1]
    hw = cw * want[3]
    for g in range(4):
        assert np.all(np.abs(got[g] - want[g]) <= bound(S[g], sum(ks), 2))
    np.testing.assert_allclose(c, cw, rtol=0, atol=3e-5)
    np.testing.assert_allclose(h, hw, rtol=0, atol=3e-5)


def test_deterministic(gpu):
    rng = np.random.default_rng(3)
    A = [rng.uniform(-1, 1, (64, k)).astype(np.float32) for k in (128, 1024)]
    B = [[rng.uniform(-1, 1, (k, 512)).astype(np.float32) for k in (128, 1024)] for _ in range(4)]
    first = run_grouped(gpu, A, B, None, [0] * 4, 0, 2)
    again = run_grouped(gpu, A, B, None, [0] * 4, 0, 2)
    for x, y in zip(first, again):
        assert np.array_equal(x, y)


def test_sequence_launch_is_bit_identical_to_per_step_launches(gpu):
    """tcr_gemm_grouped_seq_*: 12 chained LSTM gate steps in ONE resident launch (grid barrier between steps) give the bits of 12
    separate tcr_gemm_grouped launches, on repeated launches too (the barrier counters reset themselves); tools/rnn_seq_check.py
    is the full-size (128 steps, hidden 1024) form with timings."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for prec in ("1", "2"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "rnn_seq_check.py"), "12", prec, "48", "32", "256"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and '"bit_identical": true' in r.stdout, (r.stdout[-600:], r.stderr[-600:])
