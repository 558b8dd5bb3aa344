"""CPU tests: the oracle pinned to an INDEPENDENT ONNX runtime.

tests/golden/cv2dnn_*.npz hold tensors that OpenCV DNN 4.13 (third-party code) computed from the reference's own
ONNX files (tests/golden/make_golden_cv2dnn.py cuts static-shape sub-models with oracle/onnx_subgraph.py; node and
weight bytes are the reference's, verbatim).  Here
  * the committed interpreter goldens (tests/golden/sp_*.npz, lg_*.npz) and
  * the readable restatements (oracle/superpoint_ref.py, oracle/lightglue_ref.py)
are checked against those tensors; when /root/reference and cv2 are both present the fixtures are also regenerated
and compared, so a stale fixture cannot hide a disagreement.  tests/test_gpu_parity.py compares the CUDA path with the
same fixtures on the B200 box.
"""
import os

import numpy as np
import pytest
import torch

from oracle import lightglue_ref, superpoint_ref, synth
from tests import parity

REF_ONNX = "/root/reference/onnxmodel"


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["sp_640x480_seed0", "sp_752x480_seed100_a"])
def test_interpreter_goldens_agree_with_cv2dnn(golden_dir, name):
    g, c = _load(golden_dir, name), _load(golden_dir, "cv2dnn_" + name)
    assert np.array_equal(g["heat_rows"], c["heat_rows"])
    assert np.abs(g["heat"] - c["heat"]).max() <= 5e-6            # measured 1.2e-6 (fp32 summation order only)
    assert np.abs(g["dense_desc_px"] - c["dense_desc_px"]).max() <= 5e-6


@pytest.mark.parametrize("name,img", [("sp_640x480_seed0", lambda: synth.frame(0, 480, 640)),
                                      ("sp_752x480_seed100_a", lambda: synth.frame_pair(100, 480, 752)[0])])
def test_superpoint_restatement_vs_cv2dnn(golden_dir, name, img):
    c = _load(golden_dir, "cv2dnn_" + name)
    taps = {}
    sp = superpoint_ref.SuperPointRef()
    k, s, d = sp(img(), taps)
    heat = taps["heatmap"][0].numpy()
    assert np.abs(heat[c["heat_rows"]] - c["heat"]).max() <= parity.HEAT_ATOL
    assert np.abs(taps["dense_desc"][0].numpy()[:, ::8, ::8] - c["dense_desc_px"]).max() <= parity.DESC_ATOL
    # first NMS stage (MaxPool 9x9 with -inf padding, Equal, Cast, MaxPool, Greater, Where) as cv2.dnn ran it:
    # supp_scores = where(maxpool(float(s == maxpool(s))) > 0, 0, s)
    s0 = torch.from_numpy(heat)[None, None]
    mx = s0 == torch.nn.functional.max_pool2d(s0, 9, 1, 4)
    supp = torch.nn.functional.max_pool2d(mx.float(), 9, 1, 4) > 0
    nms1 = torch.where(supp, torch.zeros_like(s0), s0)[0, 0].numpy()
    # the two heat-maps differ by ~1e-6, so a pixel that is a 9x9 maximum in one map can lose to a near-tie in the other:
    # compare where both agree on the suppression mask, and require the masks to agree almost everywhere
    got, want = nms1[c["heat_rows"]], c["nms1"]
    same = (got != 0) == (want != 0)
    assert same.mean() > 0.9995
    assert np.abs(got[same] - want[same]).max() <= parity.HEAT_ATOL


@pytest.mark.parametrize("n", [256, 512])
def test_lightglue_restatement_and_golden_vs_cv2dnn(golden_dir, n):
    c, g = _load(golden_dir, f"cv2dnn_lg_synth_n{n}"), _load(golden_dir, f"lg_synth_n{n}")
    # interpreter golden match list == the list derived from cv2.dnn's log-assignment matrix
    r = parity.compare_matches(c["matches"], c["mscores"], g["matches"], g["mscores"])
    assert r["only_ref"] == 0 and r["only_tst"] == 0 and r["mscore_maxabs"] < 1e-3
    # restatement: log-assignment matrix and matches
    k0, k1, d0, d1, _ = synth.lightglue_inputs(n, 200 + n)
    taps = {}
    lg = lightglue_ref.LightGlueRef()
    m, ms = lg(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1, taps=taps)
    S = taps["S"].numpy().reshape(n, n)
    assert np.abs(S[c["S_rows"]] - c["S"]).max() <= 5e-3 * max(1.0, np.abs(c["S"]).max() / 100)   # measured 1.1e-3 on |S| <= 107
    parity.compare_matches(c["matches"], c["mscores"], m.numpy(), ms.numpy())


@pytest.mark.skipif(not os.path.isdir(REF_ONNX), reason="reference not mounted")
def test_fixtures_regenerate_from_reference_with_cv2dnn(golden_dir):
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "dnn"):
        pytest.skip("cv2 without dnn")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_cv2dnn", os.path.join(golden_dir, "make_golden_cv2dnn.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    c = _load(golden_dir, "cv2dnn_sp_640x480_seed0")
    heat, dense, nms1 = mk.cv2dnn_superpoint(synth.frame(0, 480, 640))
    assert np.abs(heat[c["heat_rows"]] - c["heat"]).max() <= 1e-6
    assert np.abs(dense[:, ::8, ::8] - c["dense_desc_px"]).max() <= 1e-6
    assert abs(int((nms1 != 0).sum()) - int(c["nms1_nonzero"])) <= 2
    # ... and the literal interpreter against cv2.dnn on the full tensors, not only the stored rows
    from oracle import onnx_interp
    it = onnx_interp.Interpreter(os.path.join(REF_ONNX, "superpoint.onnx"))
    x = torch.from_numpy(synth.frame(0, 480, 640).astype(np.float32) * np.float32(1.0 / 255.0))[None, None]
    o = it.run({"image": x}, keep=["/Reshape_1_output_0", "/Div_output_0"])
    assert np.abs(o["/Reshape_1_output_0"][0].numpy() - heat).max() <= 5e-6
    assert np.abs(o["/Div_output_0"][0].numpy() - dense).max() <= 5e-6
    cl = _load(golden_dir, "cv2dnn_lg_synth_n256")
    k0, k1, d0, d1, _ = synth.lightglue_inputs(256, 456)
    S = mk.cv2dnn_lightglue_S(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1)
    assert np.abs(S[cl["S_rows"]] - cl["S"]).max() <= 1e-4
    m, ms = mk.matches_from_S(S)
    assert np.array_equal(m, cl["matches"])
