"""Minimal ONNX (protobuf wire-format) reader -- TEST INFRASTRUCTURE ONLY.

The reference keeps its whole arithmetic in two ONNX graphs
(/root/reference/onnxmodel/superpoint.onnx, lightglue_sim.onnx; loaded at
src/Extractors/superpoint_onnx.cc:35 and src/Matchers/lightglue_onnx.cpp:45).
Neither `onnx` nor `onnxruntime` exists in this image, so this module walks the
protobuf wire format directly.  Field numbers follow onnx.proto3:

  ModelProto   {7: graph}
  GraphProto   {1: node*, 5: initializer*, 11: input*, 12: output*}
  NodeProto    {1: input*, 2: output*, 3: name, 4: op_type, 5: attribute*}
  AttributeProto {1: name, 2: f, 3: i, 4: s, 5: t, 7: floats, 8: ints, 20: type}
  TensorProto  {1: dims*, 2: data_type, 4: float_data, 7: int64_data, 8: name, 9: raw_data}
  ValueInfoProto {1: name}

Only `tests/`, `tools/pack_weights.py` and the golden-vector generator use it.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64,
           9: np.bool_, 10: np.float16, 11: np.float64}


def _varint(buf: bytes, pos: int):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt):
    if wt == 0:
        return [_signed(val)]
    out = []
    pos = 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_signed(v))
    return out


def _tensor(buf: bytes):
    dims, dtype, name, raw = [], 1, "", None
    floats, int64s, int32s = [], [], []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dims += _packed_varints(val, wt)
        elif fno == 2:
            dtype = val
        elif fno == 4:
            if wt == 5:
                floats.append(struct.unpack("<f", val)[0])
            else:
                floats += list(np.frombuffer(val, dtype="<f4"))
        elif fno == 5:
            int32s += _packed_varints(val, wt)
        elif fno == 7:
            int64s += _packed_varints(val, wt)
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = bytes(val)
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dt).newbyteorder("<")).astype(np_dt)
    elif floats:
        arr = np.asarray(floats, dtype=np_dt)
    elif int64s:
        arr = np.asarray(int64s, dtype=np_dt)
    elif int32s:
        arr = np.asarray(int32s, dtype=np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    arr = arr.reshape(dims) if dims else (arr.reshape(()) if arr.size == 1 else arr)
    return name, arr


def _attribute(buf: bytes):
    name, value = "", None
    for fno, wt, val in _fields(buf):
        if fno == 1:
            name = val.decode()
        elif fno == 2:
            value = struct.unpack("<f", val)[0]
        elif fno == 3:
            value = _signed(val)
        elif fno == 4:
            value = val.decode(errors="replace")
        elif fno == 5:
            value = _tensor(val)[1]
        elif fno == 7:
            if value is None:
                value = []
            if wt == 5:
                value.append(struct.unpack("<f", val)[0])
            else:
                value += list(np.frombuffer(val, dtype="<f4"))
        elif fno == 8:
            if value is None:
                value = []
            value += _packed_varints(val, wt)
    return name, value


@dataclass
class Node:
    op: str
    name: str
    inputs: list
    outputs: list
    attrs: dict = field(default_factory=dict)


@dataclass
class Graph:
    nodes: list
    initializers: dict
    inputs: list
    outputs: list


def _node(buf: bytes) -> Node:
    ins, outs, name, op, attrs = [], [], "", "", {}
    for fno, wt, val in _fields(buf):
        if fno == 1:
            ins.append(val.decode())
        elif fno == 2:
            outs.append(val.decode())
        elif fno == 3:
            name = val.decode()
        elif fno == 4:
            op = val.decode()
        elif fno == 5:
            k, v = _attribute(val)
            attrs[k] = v
    return Node(op, name, ins, outs, attrs)


def _value_info_name(buf: bytes) -> str:
    for fno, wt, val in _fields(buf):
        if fno == 1:
            return val.decode()
    return ""


def load(path: str) -> Graph:
    with open(path, "rb") as f:
        model = f.read()
    graph_buf = None
    for fno, wt, val in _fields(model):
        if fno == 7:
            graph_buf = val
    if graph_buf is None:
        raise ValueError("no graph in model")
    nodes, inits, ins, outs = [], {}, [], []
    for fno, wt, val in _fields(graph_buf):
        if fno == 1:
            nodes.append(_node(val))
        elif fno == 5:
            name, arr = _tensor(val)
            inits[name] = arr
        elif fno == 11:
            ins.append(_value_info_name(val))
        elif fno == 12:
            outs.append(_value_info_name(val))
    ins = [i for i in ins if i not in inits]
    return Graph(nodes, inits, ins, outs)


if __name__ == "__main__":
    import sys
    from collections import Counter
    g = load(sys.argv[1])
    print("inputs", g.inputs, "outputs", g.outputs)
    print("nodes", len(g.nodes), "initializers", len(g.initializers),
          "params", sum(a.size for a in g.initializers.values()))
    print(Counter(n.op for n in g.nodes))
