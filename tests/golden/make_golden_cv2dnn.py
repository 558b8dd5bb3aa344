#!/usr/bin/env python
"""Golden vectors from an INDEPENDENT runtime: the reference's own ONNX files executed by OpenCV DNN.

ONNXRuntime (the reference's runtime) is absent here and on the GPU box (profiles/r02_ort_probe.txt), but
opencv-python-headless 4.13 ships cv2.dnn, a third-party ONNX executor nobody in this repository wrote.  It cannot
import the full graphs (dynamic shapes, And / NonZero / ScatterND / TopK), so oracle/onnx_subgraph.py cuts static-shape
sub-models out of the reference's files -- node and weight bytes copied verbatim -- that end where cv2.dnn stops:

  superpoint.onnx     image -> /Reshape_1_output_0  (heat-map after softmax-65 + depth-to-space; nodes 1-57)
                            -> /Div_output_0        (L2-normalised dense descriptors; nodes 399-413)
                            -> /Where_output_0      (scores with the first NMS suppression applied; nodes 58-68)
  lightglue_sim.onnx  kpts/desc -> /log_assignment.8/Add_2_output_0  (the N0 x N1 log-assignment matrix after all 9
                            layers: every MatMul / Softmax / LayerNorm / Erf of the graph; nodes 0-1501)
What cv2.dnn cannot run is integer / comparison work only (NMS iterations 2-3, border, NonZero; TopK, mutual check,
filter): this script applies those steps in numpy to cv2.dnn's tensors so that the fixtures also hold match lists.

    python tests/golden/make_golden_cv2dnn.py        (needs /root/reference and cv2; Winograd disabled: plain fp32)

Outputs: cv2dnn_sp_640x480_seed0.npz, cv2dnn_sp_752x480_seed100_a.npz, cv2dnn_lg_synth_n{256,512}.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import lightglue_ref, onnx_subgraph, synth  # noqa: E402

REF = os.environ.get("ROVER_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
HEAT_STRIDE = 16
SP_OUTS = ["/Reshape_1_output_0", "/Div_output_0", "/Where_output_0"]
LG_OUT = "/log_assignment.8/Add_2_output_0"


def cv2dnn_superpoint(img, tmp="/tmp/rfe_cv2dnn_sp.onnx"):
    """u8 image -> (heat [H,W], dense [256,H/8,W/8], nms1 [H,W]) computed by cv2.dnn from the reference's graph."""
    import cv2
    h, w = img.shape
    shapes = {"/Reshape_1_output_0": [1, h, w], "/Div_output_0": [1, 256, h // 8, w // 8], "/Where_output_0": [1, 1, h, w]}
    onnx_subgraph.cut(os.path.join(REF, "onnxmodel", "superpoint.onnx"), {"image": (1, [1, 1, h, w])}, shapes, tmp)
    net = cv2.dnn.readNetFromONNX(tmp)
    net.enableWinograd(False)
    x = (img.astype(np.float32) * np.float32(1.0 / 255.0)).reshape(1, 1, h, w)     # transform.cpp:3-17
    net.setInput(x, "image")
    heat, dense, nms1 = net.forward(SP_OUTS)
    return heat.reshape(h, w), dense.reshape(256, h // 8, w // 8), nms1.reshape(h, w)


def cv2dnn_lightglue_S(kn0, kn1, d0, d1, tmp="/tmp/rfe_cv2dnn_lg.onnx"):
    """Normalised keypoints + descriptors -> log-assignment matrix [N0, N1] computed by cv2.dnn."""
    import cv2
    n0, n1 = len(kn0), len(kn1)
    feeds = {"kpts0": np.asarray(kn0, np.float32).reshape(1, n0, 2), "kpts1": np.asarray(kn1, np.float32).reshape(1, n1, 2),
             "desc0": np.asarray(d0, np.float32).reshape(1, n0, 256), "desc1": np.asarray(d1, np.float32).reshape(1, n1, 256)}
    onnx_subgraph.cut(os.path.join(REF, "onnxmodel", "lightglue_sim.onnx"), {k: (1, list(v.shape)) for k, v in feeds.items()},
                      {LG_OUT: [1, n0, n1]}, tmp)
    net = cv2.dnn.readNetFromONNX(tmp)
    net.enableWinograd(False)
    for k, v in feeds.items():
        net.setInput(v, k)
    return net.forward(LG_OUT).reshape(n0, n1)


def matches_from_S(S):
    """lightglue_sim.onnx nodes 1502-1525 in numpy: TopK(k=1) both ways (lowest index on ties), mutual check, exp, > 0.1."""
    m0, m1 = S.argmax(1), S.argmax(0)
    mutual = m1[m0] == np.arange(S.shape[0])
    ms = np.where(mutual, np.exp(S.max(1)), np.float32(0)).astype(np.float32)
    idx = np.nonzero(ms > np.float32(0.1))[0]
    return np.stack([idx, m0[idx]], 1).astype(np.int32), ms[idx]


def main():
    for name, img in (("cv2dnn_sp_640x480_seed0.npz", synth.frame(0, 480, 640)),
                      ("cv2dnn_sp_752x480_seed100_a.npz", synth.frame_pair(100, 480, 752)[0])):
        heat, dense, nms1 = cv2dnn_superpoint(img)
        rows = np.arange(0, heat.shape[0], HEAT_STRIDE, dtype=np.int32)
        np.savez_compressed(os.path.join(OUT, name), heat_rows=rows, heat=heat[rows], dense_desc_px=dense[:, ::8, ::8].copy(),
                            nms1=nms1[rows], nms1_nonzero=np.int64((nms1 != 0).sum()))
        print(name, "heat max", heat.max(), "nms1 nonzero", int((nms1 != 0).sum()))
    for n in (256, 512):
        k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 200 + n)
        S = cv2dnn_lightglue_S(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1)
        m, ms = matches_from_S(S)
        step = 1 if n == 256 else 4
        np.savez_compressed(os.path.join(OUT, f"cv2dnn_lg_synth_n{n}.npz"), S_rows=np.arange(0, n, step, dtype=np.int32),
                            S=S[::step].copy(), matches=m, mscores=ms)
        print(n, "matches", len(m))
    for f in sorted(os.listdir(OUT)):
        if f.startswith("cv2dnn"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
