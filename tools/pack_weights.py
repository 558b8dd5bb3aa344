#!/usr/bin/env python
"""Repack the reference's ONNX initialisers into the flat weight blob the library loads.

Input : /root/reference/onnxmodel/superpoint.onnx, lightglue_sim.onnx (paths hard-coded by the
        reference at src/Extractors/SPextractor.cc:93 and src/Matchers/lightglue_onnx.cpp:38).
Output: weights/rover_fe.rfw  ("RFW1" container, see rover_slam_b200/csrc/weights.h)

Layouts written (all fp32, little endian) -- chosen for the kernels, not the ONNX order:
  conv weights  : O,kh,kw,I  ("OHWI"/KRSC; the K axis of the implicit GEMM is contiguous)
  linear weights: [out, in]  (ONNX MatMul initialisers are [in, out]; transposed here so that
                              K is contiguous = the K-major B operand of tcgen05.mma)
Names: sp.<conv>.{w,b};  lg.posenc.w;  lg.l<i>.{self,cross}.<linear>.{w,b}; lg.l<i>.*.ln.{w,b};
       lg.final_proj.{w,b}; lg.matchability.{w,b}

Run in the build container only (the GPU box has no /root/reference); the blob is committed.
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import onnx_reader  # noqa: E402

REF = os.environ.get("ROVER_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "weights", "rover_fe.rfw")


def superpoint_tensors(path):
    g = onnx_reader.load(path)
    out = {}
    for name in ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
                 "convPa", "convPb", "convDa", "convDb"]:
        w = g.initializers[name + ".weight"]           # OIHW
        out[f"sp.{name}.w"] = np.ascontiguousarray(w.transpose(0, 2, 3, 1))   # OHWI
        out[f"sp.{name}.b"] = g.initializers[name + ".bias"].copy()
    return out


def lightglue_tensors(path):
    g = onnx_reader.load(path)
    init = g.initializers
    consumers = {}
    for n in g.nodes:
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)
    out = {}
    seen = {}
    for n in g.nodes:
        if n.op != "MatMul" or n.inputs[1] not in init:
            continue
        wname = n.inputs[1]
        w = init[wname]
        nxt = consumers.get(n.outputs[0], [])
        bias = None
        for c in nxt:
            if c.op == "Add":
                for i in c.inputs:
                    if i in init and init[i].ndim == 1:
                        bias = i
        if bias is None:
            assert w.shape == (2, 32), (wname, w.shape)
            canon = "lg.posenc"
        else:
            b = bias[:-len(".bias")]
            if b.startswith("transformers."):
                _, li, blk, *rest = b.split(".")
                blk = {"self_attn": "self", "cross_attn": "cross"}[blk]
                lin = "".join(rest).lower()            # Wqkv->wqkv, ffn.0 -> ffn0
                canon = f"lg.l{li}.{blk}.{lin}"
            else:
                canon = "lg." + b.split(".")[-1]       # log_assignment.8.final_proj -> final_proj
        wt = np.ascontiguousarray(w.T)
        if canon in seen:
            assert np.array_equal(seen[canon], wt), f"shared weight {canon} differs between uses"
            continue
        seen[canon] = wt
        out[canon + ".w"] = wt
        if bias is not None:
            out[canon + ".b"] = init[bias].copy()
    for k, v in init.items():
        if ".ffn.1." in k:                              # LayerNorm affine
            _, li, blk, _, _, kind = k.split(".")
            blk = {"self_attn": "self", "cross_attn": "cross"}[blk]
            out[f"lg.l{li}.{blk}.ln.{'w' if kind == 'weight' else 'b'}"] = v.copy()
    return out


def write_blob(tensors: dict, path: str):
    names = sorted(tensors)
    header = 16 + 128 * len(names)
    off = (header + 255) // 256 * 256
    table = []
    for nme in names:
        a = np.ascontiguousarray(tensors[nme], dtype=np.float32)
        dims = list(a.shape) + [1] * (4 - a.ndim)
        table.append((nme, a.ndim, dims, off, a.nbytes, a))
        off = (off + a.nbytes + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(struct.pack("<4sIQ", b"RFW1", len(names), off))
        for nme, nd, dims, o, nb, _ in table:
            f.write(struct.pack("<80sI4IQQ", nme.encode(), nd, *dims, o, nb) + b"\0" * (128 - 80 - 4 - 16 - 16))
        for nme, nd, dims, o, nb, a in table:
            f.seek(o)
            f.write(a.tobytes())
        f.truncate(off)
    return off


def main():
    t = {}
    t.update(superpoint_tensors(os.path.join(REF, "onnxmodel", "superpoint.onnx")))
    t.update(lightglue_tensors(os.path.join(REF, "onnxmodel", "lightglue_sim.onnx")))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    n = write_blob(t, OUT)
    print(f"wrote {OUT}: {len(t)} tensors, {n} bytes, {sum(v.size for v in t.values())} params")
    for k in sorted(t)[:40]:
        print("  ", k, t[k].shape)


if __name__ == "__main__":
    main()
