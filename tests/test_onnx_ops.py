"""Pins the operator semantics of the literal ONNX interpreter (oracle/onnx_interp.py -- the normative oracle, it executes the
reference's own graphs) to the ONNX operator specification: the worked examples printed in the spec (Operators.md) for the
index-manipulating ops, and independent numpy restatements of the spec's formulas for the arithmetic ones.  ONNXRuntime is
not installable here, so this is the closest pin available for "what ONNXRuntime-CPU computes" at
src/Extractors/superpoint_onnx.cc:133-136 / src/Matchers/lightglue_onnx.cpp:210-214: every operator below occurs in
superpoint.onnx or lightglue_sim.onnx (SURVEY.md 8(c) lists the 36 + 28 op types)."""
import types

import numpy as np
import pytest
import torch

from oracle import onnx_interp


@pytest.fixture(scope="module")
def it():
    i = onnx_interp.Interpreter.__new__(onnx_interp.Interpreter)      # no graph: the op_* methods only need self.dtype
    i.dtype = torch.float32
    return i


def node(outputs=1, **attrs):
    return types.SimpleNamespace(attrs=attrs, outputs=["o%d" % k for k in range(outputs)], inputs=[])


def T(x, dtype=None):
    return torch.tensor(np.asarray(x, dtype=dtype))


def test_nonzero_spec_example(it):
    out = it.op_NonZero(node(), T([[1, 0], [1, 1]], np.float32))
    assert out.dtype == torch.int64 and out.tolist() == [[0, 1, 1], [0, 0, 1]]      # row-major order of the hits


def test_scatternd_spec_examples(it):
    out = it.op_ScatterND(node(), T([1, 2, 3, 4, 5, 6, 7, 8], np.float32), T([[4], [3], [1], [7]], np.int64),
                          T([9, 10, 11, 12], np.float32))
    assert out.tolist() == [1, 11, 3, 10, 9, 6, 7, 12]
    data = np.arange(64, dtype=np.float32).reshape(4, 4, 4)
    upd = -np.arange(32, dtype=np.float32).reshape(2, 4, 4)
    out = it.op_ScatterND(node(), T(data), T([[0], [2]], np.int64), T(upd))           # slice updates (spec example 2 shape)
    want = data.copy()
    want[0], want[2] = upd[0], upd[1]
    assert np.array_equal(out.numpy(), want)
    out = it.op_ScatterND(node(), torch.zeros(2, 3), T([[0, 1], [1, 2]], np.int64), T([5, 7], np.float32))
    assert out.tolist() == [[0, 5, 0], [0, 0, 7]]


def test_gathernd_and_gather_spec_examples(it):
    data = T([[0, 1], [2, 3]], np.float32)
    assert it.op_GatherND(node(), data, T([[0, 0], [1, 1]], np.int64)).tolist() == [0, 3]
    assert it.op_GatherND(node(), data, T([[1], [0]], np.int64)).tolist() == [[2, 3], [0, 1]]
    d3 = T([[[0, 1], [2, 3]], [[4, 5], [6, 7]]], np.float32)
    assert it.op_GatherND(node(), d3, T([[0, 1], [1, 0]], np.int64)).tolist() == [[2, 3], [4, 5]]
    # Gather, axis 0 / axis 1 examples of the spec
    d = T([[1.0, 1.2], [2.3, 3.4], [4.5, 5.7]], np.float32)
    assert np.allclose(it.op_Gather(node(axis=0), d, T([[0, 1], [1, 2]], np.int64)).numpy(),
                       [[[1.0, 1.2], [2.3, 3.4]], [[2.3, 3.4], [4.5, 5.7]]])
    d = T([[1.0, 1.2, 1.9], [2.3, 3.4, 3.9], [4.5, 5.7, 5.9]], np.float32)
    assert np.allclose(it.op_Gather(node(axis=1), d, T([[0, 2]], np.int64)).numpy(), [[[1.0, 1.9]], [[2.3, 3.9]], [[4.5, 5.9]]])
    assert it.op_Gather(node(axis=0), T([10, 20, 30], np.int64), T(-1, np.int64)).item() == 30     # negative index, scalar
    # GatherElements, spec example 1
    out = it.op_GatherElements(node(axis=1), T([[1, 2], [3, 4]], np.float32), T([[0, 0], [1, 0]], np.int64))
    assert out.tolist() == [[1, 1], [4, 3]]


def test_topk_spec_example_and_tie_break(it):
    x = T(np.arange(12, dtype=np.float32).reshape(3, 4))
    v, i = it.op_TopK(node(2, axis=1, largest=1), x, T([3], np.int64))
    assert v.tolist() == [[3, 2, 1], [7, 6, 5], [11, 10, 9]] and i.tolist() == [[3, 2, 1]] * 3
    # the matcher's TopK(k=1): "if two elements are equal, the lower-index element appears first" (spec)
    v, i = it.op_TopK(node(2, axis=-1, largest=1), T([[1, 5, 5, 2], [7, 7, 7, 7]], np.float32), T([1], np.int64))
    assert v.tolist() == [[5], [7]] and i.tolist() == [[1], [0]]
    v, i = it.op_TopK(node(2, axis=0, largest=1), T([[1, 9], [4, 9], [4, 2]], np.float32), T([1], np.int64))
    assert v.tolist() == [[4, 9]] and i.tolist() == [[1, 0]]


def test_slice_spec_examples(it):
    d = T([[1, 2, 3, 4], [5, 6, 7, 8]], np.float32)
    out = it.op_Slice(node(), d, T([1, 0], np.int64), T([2, 3], np.int64), T([0, 1], np.int64), T([1, 2], np.int64))
    assert out.tolist() == [[5, 7]]
    out = it.op_Slice(node(), d, T([0, 1], np.int64), T([-1, 1000], np.int64))
    assert out.tolist() == [[2, 3, 4]]
    x = T(np.arange(10, dtype=np.float32))
    out = it.op_Slice(node(), x, T([-1], np.int64), T([-(2 ** 63) + 1], np.int64), T([0], np.int64), T([-1], np.int64))
    assert out.tolist() == list(range(9, -1, -1))                                        # full reversal idiom
    out = it.op_Slice(node(), x, T([7], np.int64), T([2], np.int64), T([0], np.int64), T([-2], np.int64))
    assert out.tolist() == [7, 5, 3]


def test_shape_plumbing_ops(it):
    assert it.op_Range(node(), T(1, np.int64), T(7, np.int64), T(2, np.int64)).tolist() == [1, 3, 5]
    x = T([[1], [2], [3]], np.float32)
    assert it.op_Expand(node(), x, T([2, 1, 6], np.int64)).shape == (2, 3, 6)          # spec "dim_changed" example
    assert it.op_Expand(node(), x, T([3, 4], np.int64)).tolist() == [[1] * 4, [2] * 4, [3] * 4]
    c = it.op_ConstantOfShape(node(value=np.array([5], np.int64)), T([2, 3], np.int64))
    assert c.dtype == torch.int64 and c.tolist() == [[5, 5, 5], [5, 5, 5]]
    assert it.op_ConstantOfShape(node(), T([4], np.int64)).tolist() == [0.0] * 4       # default: float32 zeros
    y = T(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    assert it.op_Flatten(node(axis=2), y).shape == (6, 4) and it.op_Flatten(node(axis=0), y).shape == (1, 24)
    assert it.op_Reshape(node(), y, T([0, -1], np.int64)).shape == (2, 12)             # 0 copies the input dimension
    assert it.op_Unsqueeze(node(), T([1.0, 2.0]), T([0, 2], np.int64)).shape == (1, 2, 1)
    assert it.op_Squeeze(node(), torch.zeros(1, 3, 1, 2), T([0, 2], np.int64)).shape == (3, 2)
    a, b = it.op_Split(node(2, axis=1), T(np.arange(12, dtype=np.float32).reshape(2, 6)), T([2, 4], np.int64))
    assert a.shape == (2, 2) and b.tolist() == [[2, 3, 4, 5], [8, 9, 10, 11]]
    assert it.op_Transpose(node(perm=[2, 0, 1]), y).shape == (4, 2, 3)
    assert it.op_Concat(node(axis=1), T([[1, 2]]), T([[3]])).tolist() == [[1, 2, 3]]
    assert it.op_Shape(node(), y).tolist() == [2, 3, 4]


def test_elementwise_and_logic_ops(it):
    a = T([-1.5, 0.0, 2.5, 7.9], np.float32)
    assert it.op_Cast(node(to=7), a).tolist() == [-1, 0, 2, 7]                          # float -> int64 truncates toward zero
    assert it.op_Cast(node(to=9), a).tolist() == [True, False, True, True]
    assert it.op_Clip(node(), a, T(0.0, np.float32), T(3.0, np.float32)).tolist() == [0.0, 0.0, 2.5, 3.0]
    assert it.op_Clip(node(), a, T(1e-12, np.float32)).tolist()[1] == pytest.approx(1e-12)    # L2-normalise: clip(min) only
    c = T([True, False, True, False])
    assert np.array_equal(it.op_Where(node(), c, a, -a).numpy(), np.array([-1.5, -0.0, 2.5, -7.9], np.float32))
    assert it.op_And(node(), c, ~c).any().item() is False and it.op_Or(node(), c, ~c).all().item() is True
    assert it.op_Not(node(), c).tolist() == [False, True, False, True]
    assert it.op_Equal(node(), a, T([-1.5, 1.0, 2.5, 0.0], np.float32)).tolist() == [True, False, True, False]
    assert it.op_Greater(node(), a, torch.zeros(4)).tolist() == [False, False, True, True]
    x = np.linspace(-3, 3, 13).astype(np.float32)
    from math import erf
    assert np.allclose(it.op_Erf(node(), T(x)).numpy(), [erf(float(v)) for v in x], atol=1e-6)
    assert np.allclose(it.op_Sigmoid(node(), T(x)).numpy(), 1 / (1 + np.exp(-x.astype(np.float64))), atol=1e-6)
    assert np.allclose(it.op_Pow(node(), T(x), T(2.0, np.float32)).numpy(), x * x)
    assert np.allclose(it.op_ReduceSum(node(keepdims=1), T(np.ones((2, 3), np.float32)), T([1], np.int64)).numpy(), [[3], [3]])
    assert it.op_ReduceSum(node(keepdims=0), T(np.ones((2, 3), np.float32)), T([-1], np.int64)).shape == (2,)
    m = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert np.allclose(it.op_MatMul(node(), T(np.stack([m, 2 * m])), T(m.T)).numpy(), np.stack([m @ m.T, 2 * m @ m.T]))


def test_softmax_and_layernorm_follow_the_spec_formulas(it):
    rng = np.random.RandomState(0)
    x = rng.randn(2, 5, 7).astype(np.float32) * 3
    for axis in (1, -1):
        e = np.exp(x.astype(np.float64) - x.max(axis=axis, keepdims=True))
        sm = e / e.sum(axis=axis, keepdims=True)                                       # opset >= 13: along `axis` only
        assert np.allclose(it.op_Softmax(node(axis=axis), T(x)).numpy(), sm, atol=1e-6)
        assert np.allclose(it.op_LogSoftmax(node(axis=axis), T(x)).numpy(), np.log(sm), atol=1e-5)
    w, b = rng.randn(7).astype(np.float32), rng.randn(7).astype(np.float32)
    x64 = x.astype(np.float64)
    mean = x64.mean(-1, keepdims=True)
    var = ((x64 - mean) ** 2).mean(-1, keepdims=True)                                  # population variance, eps inside the sqrt
    want = (x64 - mean) / np.sqrt(var + 1e-5) * w + b
    assert np.allclose(it.op_LayerNormalization(node(axis=-1, epsilon=1e-5), T(x), T(w), T(b)).numpy(), want, atol=1e-5)


def test_conv_and_maxpool_against_brute_force(it):
    rng = np.random.RandomState(1)
    x = rng.randn(1, 2, 6, 7).astype(np.float32)
    w = rng.randn(3, 2, 3, 3).astype(np.float32)
    b = rng.randn(3).astype(np.float32)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))               # Conv pads with zeros
    want = np.zeros((1, 3, 6, 7))
    for o in range(3):
        for yy in range(6):
            for xx in range(7):
                want[0, o, yy, xx] = (xp[0, :, yy:yy + 3, xx:xx + 3] * w[o]).sum() + b[o]     # cross-correlation, no flip
    got = it.op_Conv(node(pads=[1, 1, 1, 1], kernel_shape=[3, 3], strides=[1, 1]), T(x), T(w), T(b)).numpy()
    assert np.allclose(got, want, atol=1e-5)
    # MaxPool pads with -inf (NOT zero): matters for the NMS max-pool of scores that are all >= 0 only at the -1 border
    x = -np.abs(rng.randn(1, 1, 5, 6)).astype(np.float32) - 1.0
    xp = np.pad(x, ((0, 0), (0, 0), (4, 4), (4, 4)), constant_values=-np.inf)
    want = np.array([[xp[0, 0, yy:yy + 9, xx:xx + 9].max() for xx in range(6)] for yy in range(5)])
    got = it.op_MaxPool(node(kernel_shape=[9, 9], pads=[4, 4, 4, 4], strides=[1, 1]), T(x)).numpy()[0, 0]
    assert np.array_equal(got, want) and (got < 0).all()
    got = it.op_MaxPool(node(kernel_shape=[2, 2], strides=[2, 2]), T(x[:, :, :4, :])).numpy()[0, 0]
    assert np.array_equal(got, x[0, 0, :4].reshape(2, 2, 3, 2).max(axis=(1, 3)))


def test_gridsample_bilinear_align_corners_follows_the_spec(it):
    """GridSample-16, mode=bilinear, padding_mode=zeros, align_corners=1 (the descriptor sampler): a normalised coordinate g maps
    to pixel (g + 1) / 2 * (size - 1); grid[..., 0] is x (width), grid[..., 1] is y (height); out-of-range corners read 0."""
    rng = np.random.RandomState(2)
    x = rng.randn(1, 3, 4, 5).astype(np.float32)
    g = (rng.rand(1, 1, 9, 2).astype(np.float32) * 2.4 - 1.2)                          # some samples outside [-1, 1]
    g[0, 0, 0] = [-1.0, -1.0]                                                           # exact corners
    g[0, 0, 1] = [1.0, 1.0]
    got = it.op_GridSample(node(mode="bilinear", padding_mode="zeros", align_corners=1), T(x), T(g)).numpy()
    H, W = 4, 5
    for k in range(9):
        px, py = (g[0, 0, k, 0] + 1) / 2 * (W - 1), (g[0, 0, k, 1] + 1) / 2 * (H - 1)
        x0, y0 = int(np.floor(px)), int(np.floor(py))
        acc = np.zeros(3)
        for yy, wy in ((y0, 1 - (py - y0)), (y0 + 1, py - y0)):
            for xx, wx in ((x0, 1 - (px - x0)), (x0 + 1, px - x0)):
                if 0 <= yy < H and 0 <= xx < W:
                    acc += wy * wx * x[0, :, yy, xx]
        assert np.allclose(got[0, :, 0, k], acc, atol=1e-5), k
    assert np.allclose(got[0, :, 0, 0], x[0, :, 0, 0]) and np.allclose(got[0, :, 0, 1], x[0, :, 3, 4])
