"""Zero-copy interop: every tensor exposes __cuda_array_interface__ over its resident HBM buffer, and
EVariable.assign takes device arrays (HBM -> HBM) — the python demos keep their numpy flow, GPU pipelines skip the host."""
import numpy as np
import pytest

import tenncor_b200 as tc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_results_are_visible_to_torch_without_a_copy(gpu):
    rng = np.random.default_rng(0)
    a = tc.variable(rng.uniform(-1, 1, (6, 5, 4)).astype(np.float32), "a")
    out = tc.api.tanh(a) * 2.0
    with pytest.raises(Exception, match="no device data"):
        tc.api.exp(a).__cuda_array_interface__  # never evaluated
    out.calc()
    tc.sync()
    cai = out.__cuda_array_interface__
    assert cai["shape"] == (6, 5, 4) and cai["typestr"] == "<f4" and cai["data"][0] == out.device_ptr() and cai["version"] == 3
    t = torch.as_tensor(out, device="cuda")
    assert t.data_ptr() == out.device_ptr() and tuple(t.shape) == (6, 5, 4)
    np.testing.assert_array_equal(t.cpu().numpy(), out.data())


def test_assign_from_a_device_array(gpu):
    rng = np.random.default_rng(1)
    x = tc.EVariable([4, 7], 0, "x")
    y = tc.api.square(x)
    src = torch.as_tensor(rng.uniform(-3, 3, (4, 7)).astype(np.float32)).cuda()
    torch.cuda.synchronize()
    v0 = x.get_version()
    x.assign(src)
    assert x.get_version() > v0
    np.testing.assert_allclose(y.get(), src.cpu().numpy() ** 2, rtol=1e-6)
    other = tc.variable(rng.uniform(-1, 1, (4, 7)).astype(np.float32), "other")
    z = other + 1.0
    z.calc()
    tc.sync()
    x.assign(z)  # another tensor of this module is a device array too
    np.testing.assert_array_equal(x.data(), z.data())
    with pytest.raises(Exception, match="dtype"):
        x.assign(src.double())
    with pytest.raises(Exception, match="shaped"):
        x.assign(src[:2].contiguous())


def test_get_later_reads_every_step_one_step_behind(gpu):
    """ETensor.get_later: the copy to pinned host memory is queued behind the evaluation and the host goes on launching; result()
    returns what get() would have returned at that point, also when the variable has been overwritten since (the copy is ordered
    on the device before the next step's work)."""
    rng = np.random.default_rng(2)
    x = tc.EVariable([64, 33], 0, "x")
    w = tc.variable(rng.uniform(-1, 1, (33, 9)).astype(np.float32), "w")
    y = tc.api.reduce_sum(tc.api.tanh(tc.api.matmul(x, w)))
    batches = [rng.uniform(-1, 1, (64, 33)).astype(np.float32) for _ in range(6)]
    want = []
    for b in batches:
        x.assign(b)
        want.append(np.array(y.get()))
    pending, got = None, []
    for b in batches:
        x.assign(b)
        nxt = y.get_later()
        if pending is not None:
            assert pending.done() is False
            got.append(pending.result())
            assert pending.done() is True
        pending = nxt
    got.append(pending.result())
    got.append(pending.result())  # a second result() returns the same value
    for a, b in zip(got, want + want[-1:]):
        assert a.shape == b.shape and a.dtype == b.dtype
        np.testing.assert_array_equal(a, b)
    with pytest.raises(Exception, match="no device data"):
        tc.api.exp(x + 1.0).__cuda_array_interface__
