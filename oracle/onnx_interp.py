"""Literal ONNX-graph interpreter on torch-CPU -- TEST INFRASTRUCTURE ONLY.

Normative oracle: it executes the reference's own graphs
(/root/reference/onnxmodel/{superpoint,lightglue_sim}.onnx) node by node with the
operator semantics of ONNX opset 16/17, i.e. what ONNXRuntime-CPU computes at
src/Extractors/superpoint_onnx.cc:133-136 and src/Matchers/lightglue_onnx.cpp:210-214.
It can only run where /root/reference exists (this container); it is used to
(1) validate the readable restatement in superpoint_ref.py / lightglue_ref.py and
(2) generate the committed golden vectors under tests/golden/.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_reader

_ONNX2TORCH = {1: torch.float32, 6: torch.int32, 7: torch.int64, 9: torch.bool,
               11: torch.float64, 2: torch.uint8}


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    x = np.asarray(x)
    return torch.from_numpy(x.copy() if x.ndim else x.reshape(1).copy()).reshape(x.shape)


def _ints(t):
    return [int(v) for v in t.reshape(-1).tolist()]


class Interpreter:
    def __init__(self, path: str, dtype=torch.float32):
        self.graph = onnx_reader.load(path)
        self.dtype = dtype
        self.consts = {}
        for k, v in self.graph.initializers.items():
            t = _t(v.copy())
            if t.dtype == torch.float32 and dtype != torch.float32:
                t = t.to(dtype)
            self.consts[k] = t

    # -- execution ---------------------------------------------------------
    def run(self, feeds: dict, keep: list | None = None):
        env = dict(self.consts)
        for k, v in feeds.items():
            t = _t(v)
            if t.dtype == torch.float32 and self.dtype != torch.float32:
                t = t.to(self.dtype)
            env[k] = t
        kept = {}
        for node in self.graph.nodes:
            ins = [env[i] if i != "" else None for i in node.inputs]
            outs = getattr(self, "op_" + node.op)(node, *ins)
            if not isinstance(outs, (tuple, list)):
                outs = (outs,)
            for name, val in zip(node.outputs, outs):
                env[name] = val
                if keep and name in keep:
                    kept[name] = val
        result = {o: env[o] for o in self.graph.outputs}
        result.update(kept)
        return result

    # -- ops ----------------------------------------------------------------
    def op_Constant(self, n):
        v = n.attrs.get("value")
        if v is None:
            for k in ("value_float", "value_int", "value_floats", "value_ints"):
                if k in n.attrs:
                    v = np.asarray(n.attrs[k])
        t = _t(np.array(v, copy=True))
        if t.dtype == torch.float32 and self.dtype != torch.float32:
            t = t.to(self.dtype)
        return t

    def op_Conv(self, n, x, w, b=None):
        pads = n.attrs.get("pads", [0, 0, 0, 0])
        assert pads[0] == pads[2] and pads[1] == pads[3]
        return F.conv2d(x, w, b, stride=tuple(n.attrs.get("strides", [1, 1])),
                        padding=(pads[0], pads[1]),
                        dilation=tuple(n.attrs.get("dilations", [1, 1])),
                        groups=n.attrs.get("group", 1))

    def op_Relu(self, n, x):
        return torch.relu(x)

    def op_MaxPool(self, n, x):
        k = n.attrs["kernel_shape"]
        pads = n.attrs.get("pads", [0, 0, 0, 0])
        assert pads[0] == pads[2] and pads[1] == pads[3]
        # ONNX pads with -inf; so does torch.
        return F.max_pool2d(x, tuple(k), stride=tuple(n.attrs.get("strides", k)),
                            padding=(pads[0], pads[1]),
                            ceil_mode=bool(n.attrs.get("ceil_mode", 0)))

    def op_Softmax(self, n, x):
        return torch.softmax(x, dim=n.attrs.get("axis", -1))

    def op_LogSoftmax(self, n, x):
        return torch.log_softmax(x, dim=n.attrs.get("axis", -1))

    def op_Slice(self, n, x, starts, ends, axes=None, steps=None):
        starts, ends = _ints(starts), _ints(ends)
        axes = _ints(axes) if axes is not None else list(range(len(starts)))
        steps = _ints(steps) if steps is not None else [1] * len(starts)
        for s, e, a, st in zip(starts, ends, axes, steps):
            dim = x.shape[a]
            if st > 0:
                s = max(0, min(dim, s + dim if s < 0 else s))
                e = max(0, min(dim, e + dim if e < 0 else e))
                idx = torch.arange(s, e, st)
            else:
                s = s + dim if s < 0 else s
                s = max(-1, min(dim - 1, s))
                if e < -dim:
                    e = -1
                else:
                    e = e + dim if e < 0 else e
                    e = max(-1, min(dim - 1, e))
                idx = torch.arange(s, e, st)
            x = x.index_select(a, idx)
        return x

    def op_Shape(self, n, x):
        return torch.tensor(list(x.shape), dtype=torch.int64)

    def op_Gather(self, n, x, idx):
        axis = n.attrs.get("axis", 0)
        idx = idx.to(torch.int64)
        idx = torch.where(idx < 0, idx + x.shape[axis], idx)
        out = x.index_select(axis, idx.reshape(-1))
        shape = list(x.shape[:axis]) + list(idx.shape) + list(x.shape[axis + 1:] if axis != -1 else [])
        return out.reshape(shape)

    def op_GatherElements(self, n, x, idx):
        return torch.gather(x, n.attrs.get("axis", 0), idx.to(torch.int64))

    def op_GatherND(self, n, x, idx):
        assert n.attrs.get("batch_dims", 0) == 0
        idx = idx.to(torch.int64)
        k = idx.shape[-1]
        return x[tuple(idx[..., i] for i in range(k))]

    def op_Transpose(self, n, x):
        return x.permute(*n.attrs["perm"])

    def op_Unsqueeze(self, n, x, axes):
        for a in sorted(_ints(axes)):
            x = x.unsqueeze(a)
        return x

    def op_Squeeze(self, n, x, axes=None):
        if axes is None:
            return x.squeeze()
        for a in sorted(_ints(axes), reverse=True):
            x = x.squeeze(a)
        return x

    def op_Concat(self, n, *xs):
        return torch.cat(xs, dim=n.attrs["axis"])

    def op_Reshape(self, n, x, shape):
        shape = _ints(shape)
        shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
        return x.reshape(shape)

    def op_Flatten(self, n, x):
        a = n.attrs.get("axis", 1)
        return x.reshape(int(np.prod(x.shape[:a])) if a else 1, -1)

    def op_Mul(self, n, a, b):
        return a * b

    def op_Add(self, n, a, b):
        return a + b

    def op_Sub(self, n, a, b):
        return a - b

    def op_Div(self, n, a, b):
        if not a.is_floating_point() and not b.is_floating_point():
            return torch.div(a, b, rounding_mode="trunc")
        return a / b

    def op_Neg(self, n, a):
        return -a

    def op_Abs(self, n, a):
        return a.abs()

    def op_Pow(self, n, a, b):
        return torch.pow(a, b)

    def op_Exp(self, n, a):
        return torch.exp(a)

    def op_Log(self, n, a):
        return torch.log(a)

    def op_Erf(self, n, a):
        return torch.erf(a)

    def op_Cos(self, n, a):
        return torch.cos(a)

    def op_Sin(self, n, a):
        return torch.sin(a)

    def op_Sigmoid(self, n, a):
        return torch.sigmoid(a)

    def op_MatMul(self, n, a, b):
        return torch.matmul(a, b)

    def op_ReduceSum(self, n, x, axes=None):
        keep = bool(n.attrs.get("keepdims", 1))
        if axes is None:
            return x.sum() if not keep else x.sum().reshape([1] * x.dim())
        return x.sum(dim=_ints(axes), keepdim=keep)

    def op_Clip(self, n, x, lo=None, hi=None):
        return torch.clamp(x, min=None if lo is None else lo.item(), max=None if hi is None else hi.item())

    def op_ConstantOfShape(self, n, shape):
        v = n.attrs.get("value")
        v = _t(np.array(v, copy=True)) if v is not None else torch.zeros(1)
        if v.dtype == torch.float32 and self.dtype != torch.float32:
            v = v.to(self.dtype)
        return v.reshape(()).expand(_ints(shape)).clone() if len(_ints(shape)) else v.reshape(())

    def op_Equal(self, n, a, b):
        return a == b

    def op_Greater(self, n, a, b):
        return a > b

    def op_Not(self, n, a):
        return ~a

    def op_And(self, n, a, b):
        return a & b

    def op_Or(self, n, a, b):
        return a | b

    def op_Where(self, n, c, a, b):
        return torch.where(c, a, b)

    def op_Cast(self, n, x):
        to = _ONNX2TORCH[n.attrs["to"]]
        if to == torch.float32:
            to = self.dtype
        return x.to(to)

    def op_Expand(self, n, x, shape):
        shape = _ints(shape)
        tgt = torch.broadcast_shapes(tuple(x.shape), tuple(shape))
        return x.expand(tgt)

    def op_Range(self, n, start, limit, delta):
        return torch.arange(start.item(), limit.item(), delta.item(), dtype=start.dtype)

    def op_ScatterND(self, n, data, idx, upd):
        out = data.clone()
        idx = idx.to(torch.int64)
        k = idx.shape[-1]
        flat_idx = idx.reshape(-1, k)
        upd = upd.reshape((flat_idx.shape[0],) + tuple(data.shape[k:]))
        out[tuple(flat_idx[:, i] for i in range(k))] = upd.to(out.dtype)
        return out

    def op_NonZero(self, n, x):
        return torch.nonzero(x).t().contiguous()

    def op_Split(self, n, x, split=None):
        axis = n.attrs.get("axis", 0)
        if split is None:
            k = len(n.outputs)
            return torch.chunk(x, k, dim=axis)
        return torch.split(x, _ints(split), dim=axis)

    def op_GridSample(self, n, x, grid):
        return F.grid_sample(x, grid, mode=n.attrs.get("mode", "bilinear"),
                             padding_mode=n.attrs.get("padding_mode", "zeros"),
                             align_corners=bool(n.attrs.get("align_corners", 0)))

    def op_LayerNormalization(self, n, x, w, b=None):
        axis = n.attrs.get("axis", -1)
        assert axis in (-1, x.dim() - 1)
        return F.layer_norm(x, (x.shape[-1],), w, b, eps=n.attrs.get("epsilon", 1e-5))

    def op_TopK(self, n, x, k):
        k = int(k.reshape(-1)[0])
        if k == 1 and n.attrs.get("largest", 1):
            # ONNX breaks ties towards the lowest index; so does torch.max on CPU.
            return x.max(dim=n.attrs.get("axis", -1), keepdim=True)
        v, i = torch.topk(x, k, dim=n.attrs.get("axis", -1),
                          largest=bool(n.attrs.get("largest", 1)), sorted=True)
        return v, i
