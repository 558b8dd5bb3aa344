"""Cut a static-shape sub-model out of one of the reference's ONNX files -- TEST INFRASTRUCTURE ONLY.

Why: ONNXRuntime (the reference's runtime, README.md:22) exists neither in this container nor on the GPU box
(profiles/r02_ort_probe.txt), but OpenCV's DNN module -- an independent, third-party ONNX executor -- does
(opencv-python-headless 4.13).  cv2.dnn cannot import the full graphs (dynamic shapes, NonZero, ScatterND ...), so this
module re-serialises a SUB-graph: the original NodeProto / TensorProto bytes are copied verbatim (nothing is
re-encoded, so the weights and attributes are exactly the reference's), only the graph inputs / outputs are
rewritten with static shapes.  tests/test_oracle_cv2dnn.py runs such sub-models through cv2.dnn and compares
with oracle/onnx_interp.py: that pins the interpreter's Conv / Relu / MaxPool / MatMul / Softmax / LayerNorm semantics
(weight layout, padding, axis conventions) to a runtime nobody in this repository wrote.

Wire format: see oracle/onnx_reader.py.  ValueInfoProto {1: name, 2: TypeProto {1: Tensor {1: elem_type, 2: Shape {1: Dim {1: value}}}}}.
"""
from __future__ import annotations

from . import onnx_reader as R


def _enc_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_bytes(fno: int, payload: bytes) -> bytes:
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def _field_varint(fno: int, v: int) -> bytes:
    return _enc_varint(fno << 3) + _enc_varint(v)


def value_info(name: str, elem_type: int, shape) -> bytes:
    tensor = _field_varint(1, elem_type)
    if shape is not None:
        tensor += _field_bytes(2, b"".join(_field_bytes(1, _field_varint(1, int(d))) for d in shape))
    return _field_bytes(1, name.encode()) + _field_bytes(2, _field_bytes(1, tensor))


def cut(path: str, inputs: dict, outputs: list, out_path: str, out_elem_type: int = 1) -> list:
    """inputs: {tensor name: (onnx elem_type, static shape)}; outputs: {tensor name: static shape}.  Writes the sub-model holding
    every node needed to compute `outputs` from `inputs` and the initialisers.  Returns the op types used."""
    with open(path, "rb") as f:
        model = f.read()
    head, graph_buf = [], None
    for fno, wt, val in R._fields(model):
        if fno == 7:
            graph_buf = val
        elif wt == 0:
            head.append(_field_varint(fno, val))
        elif wt == 2:
            head.append(_field_bytes(fno, bytes(val)))
    nodes, inits = [], {}
    for fno, wt, val in R._fields(graph_buf):
        if fno == 1:
            nodes.append((R._node(val), bytes(val)))
        elif fno == 5:
            name = ""
            for f2, w2, v2 in R._fields(val):
                if f2 == 8:
                    name = v2.decode()
            inits[name] = bytes(val)
    producer = {}
    for i, (n, _) in enumerate(nodes):
        for o in n.outputs:
            producer[o] = i
    need_nodes, need_inits, stack, seen = set(), set(), list(outputs), set()
    while stack:
        t = stack.pop()
        if t in seen or t == "" or t in inputs:
            continue
        seen.add(t)
        if t in inits:
            need_inits.add(t)
            continue
        i = producer[t]
        need_nodes.add(i)
        stack.extend(nodes[i][0].inputs)
    g = b"".join(_field_bytes(1, nodes[i][1]) for i in sorted(need_nodes))
    g += _field_bytes(2, b"rover_fe_subgraph")
    g += b"".join(_field_bytes(5, inits[k]) for k in sorted(need_inits))
    g += b"".join(_field_bytes(11, value_info(k, et, shp)) for k, (et, shp) in inputs.items())
    g += b"".join(_field_bytes(12, value_info(o, out_elem_type, shp)) for o, shp in outputs.items())
    with open(out_path, "wb") as f:
        f.write(b"".join(head) + _field_bytes(7, g))
    return sorted({nodes[i][0].op for i in need_nodes})
