"""Property tests (hypothesis) for the host-side invariants everything else leans on: arena layout/bucketing, optimizer flat ≡ per-tensor,
K-of-P reduction, codec round trip."""
import numpy as np
import torch
import torch.nn as nn
from hypothesis import given, settings, strategies as st

from ps_pytorch_b200.optim import SGD, Adam
from ps_pytorch_b200.parallel.arena import ALIGN, FlatBuffer, ParamLayout
from ps_pytorch_b200.parallel.transport import reduce_k_of_p
from ps_pytorch_b200.utils.codec import g_compress, g_decompress


def _module(shapes):
    m = nn.Module()
    for i, s in enumerate(shapes):
        m.register_parameter("p%d" % i, nn.Parameter(torch.randn(*s)))
    return m


shape_st = st.one_of(st.tuples(st.integers(1, 40)), st.tuples(st.integers(1, 12), st.integers(1, 12)),
                     st.tuples(st.integers(1, 6), st.integers(1, 6), st.integers(1, 3), st.integers(1, 3)))


@settings(max_examples=40, deadline=None)
@given(st.lists(shape_st, min_size=1, max_size=12), st.integers(1, 16), st.booleans())
def test_layout_and_buckets_partition_the_arena(shapes, nb, cl):
    lay = ParamLayout.from_module(_module(shapes), channels_last=cl, chunk=256)
    assert lay.total % 256 == 0 and lay.total >= sum(lay.numels)
    for i in range(len(shapes)):
        assert lay.offsets[i] % ALIGN == 0
        if i:
            assert lay.offsets[i] >= lay.offsets[i - 1] + lay.numels[i - 1]          # no overlap, forward order
    bk = lay.buckets(nb)
    assert 1 <= len(bk) <= min(nb, len(shapes))
    assert bk[0][0] == 0 and bk[-1][1] == lay.total
    for (lo, hi, idx), nxt in zip(bk, bk[1:] + [None]):
        assert lo < hi and lo % ALIGN == 0
        if nxt is not None:
            assert hi == nxt[0]                                                       # contiguous cover of [0, total)
        for i in idx:
            assert lo <= lay.offsets[i] and lay.offsets[i] + lay.numels[i] <= hi
    assert sorted(i for _, _, idx in bk for i in idx) == list(range(len(shapes)))     # every parameter in exactly one bucket


@settings(max_examples=25, deadline=None)
@given(st.lists(shape_st, min_size=1, max_size=6), st.floats(0.001, 0.5), st.floats(0.0, 0.95), st.floats(0.0, 0.01), st.booleans(), st.integers(1, 4))
def test_sgd_flat_equals_per_tensor(shapes, lr, mom, wd, nesterov, steps):
    nesterov = nesterov and mom > 0
    torch.manual_seed(0)
    m1, m2 = _module(shapes), _module(shapes)
    m2.load_state_dict(m1.state_dict())
    lay = ParamLayout.from_module(m2, channels_last=False, chunk=256)
    master = FlatBuffer(lay, torch.float32, "cpu")
    master.bind_params(m2)
    o1 = SGD(m1.parameters(), lr=lr, momentum=mom, weight_decay=wd, nesterov=nesterov)
    o2 = SGD(m2.parameters(), lr=lr, momentum=mom, weight_decay=wd, nesterov=nesterov)
    o2.attach_flat(master)
    gflat = FlatBuffer(lay, torch.float32, "cpu")
    for _ in range(steps):
        grads = [torch.randn(*s) for s in shapes]
        o1.step([g.numpy() for g in grads])                 # reference calling convention: numpy arrays aligned with the parameters
        gflat.load_list(grads)
        o2.step(gflat.flat)
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


@settings(max_examples=15, deadline=None)
@given(st.lists(shape_st, min_size=1, max_size=4), st.floats(0.0005, 0.05), st.booleans(), st.integers(1, 3))
def test_adam_flat_equals_per_tensor(shapes, lr, amsgrad, steps):
    torch.manual_seed(1)
    m1, m2 = _module(shapes), _module(shapes)
    m2.load_state_dict(m1.state_dict())
    lay = ParamLayout.from_module(m2, channels_last=False, chunk=256)
    master = FlatBuffer(lay, torch.float32, "cpu")
    master.bind_params(m2)
    o1, o2 = Adam(m1.parameters(), lr=lr, amsgrad=amsgrad), Adam(m2.parameters(), lr=lr, amsgrad=amsgrad)
    o2.attach_flat(master)
    gflat = FlatBuffer(lay, torch.float32, "cpu")
    for _ in range(steps):
        grads = [torch.randn(*s) for s in shapes]
        o1.step(grads)
        gflat.load_list(grads)
        o2.step(gflat.flat)
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 8), st.integers(1, 8), st.booleans(), st.integers(1, 64))
def test_k_of_p_reduction(p, k, strict, n):
    k = min(k, p)
    gs = [torch.randn(n) for _ in range(p)]
    out = reduce_k_of_p(gs, k, strict)
    take = min(p, k + 1 if strict else k)
    assert torch.allclose(out, torch.stack(gs[:take]).sum(0) / k, rtol=1e-5, atol=1e-6)


@settings(max_examples=40, deadline=None)
@given(st.sampled_from([np.float32, np.float64, np.float16, np.int32, np.uint8]), st.lists(st.integers(0, 9), min_size=0, max_size=4), st.integers(0, 2 ** 31 - 1))
def test_codec_round_trip_any_array(dtype, shape, seed):
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(shape) * 5).astype(dtype)
    b = g_decompress(g_compress(a))
    assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)
