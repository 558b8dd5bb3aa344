"""GPU test of the reference-facing C++ class surface (ORB_SLAM3::SPextractor / SPmatcher and the two runner
classes) built on the C ABI: the host_driver binary follows the reference's call pattern
(Frame.cc:544-559 -> SPextractor::operator(); Tracking.cc:3465 -> SPmatcher::MatchingPoints_onnx) and its
results must equal the C-ABI path bit for bit and the oracle within tolerance."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import lightglue_ref, synth
from tests import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_spextractor_spmatcher_classes():
    from rover_slam_b200 import FrontEnd
    drv = os.path.join(ROOT, "rover_slam_b200", "host_driver")
    if not os.path.exists(drv):
        subprocess.run(["make", "-C", ROOT, "host"], check=True)
    h, w = 240, 320
    a, b = synth.frame_pair(31, h, w, shift=(7, -4))
    with tempfile.TemporaryDirectory() as td:
        a.tofile(os.path.join(td, "a.raw"))
        b.tofile(os.path.join(td, "b.raw"))
        out = os.path.join(td, "out.bin")
        env = dict(os.environ, ROVER_FE_WEIGHTS=os.path.join(ROOT, "weights", "rover_fe.rfw"))
        r = subprocess.run([drv, str(h), str(w), os.path.join(td, "a.raw"), os.path.join(td, "b.raw"), out],
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = np.fromfile(out, dtype=np.uint8)
    hdr = raw[:24].view(np.int32)
    na, nb, nmulti, m_frame, m_kp, n_ada = (int(v) for v in hdr[:6])
    off = 24
    feats = []
    for n in (na, nb):
        kp = raw[off:off + n * 12].view(np.float32).reshape(n, 3); off += n * 12
        de = raw[off:off + n * 1024].view(np.float32).reshape(n, 256); off += n * 1024
        feats.append((kp, de))
    vn_frame = raw[off:off + na * 4].view(np.int32); off += na * 4
    vn_kp = raw[off:off + na * 4].view(np.int32); off += na * 4
    kp_ada = raw[off:off + n_ada * 12].view(np.float32).reshape(n_ada, 3); off += n_ada * 12
    de_ada = raw[off:off + n_ada * 1024].view(np.float32).reshape(n_ada, 256); off += n_ada * 1024
    m_pf, m_fp, n_app_ret, n_app_ok, m_kpinf, n_vn_pf = (int(v) for v in raw[off:off + 24].view(np.int32)); off += 24
    vn_pf = raw[off:off + n_vn_pf * 4].view(np.int32); off += n_vn_pf * 4
    vn_kpinf = raw[off:off + na * 4].view(np.int32); off += na * 4
    px0 = raw[off:off + 16].view(np.float32)
    assert nmulti == 0                                   # nLevels != 1 extracts nothing, like the reference
    fe = FrontEnd(max_batch=2, max_height=h, max_width=w)
    ref = fe.extract(np.stack([a, b]))
    for (kp, de), (k, s, d) in zip(feats, ref):
        assert np.array_equal(kp[:, :2], k.astype(np.float32))
        assert np.array_equal(kp[:, 2], s)               # response = score (the reference's scores[2*i] bug waived)
        assert np.array_equal(de, d)
    # Frame overload normalises with the image size; the KeyPoint overload with the hard-coded 300 x 400
    for vn, cnt, (nh, nw) in ((vn_frame, m_frame, (h, w)), (vn_kp, m_kp, (300, 400))):
        m, ms = fe.match(ref[0][0], ref[1][0], ref[0][2], ref[1][2], nh, nw)
        exp, c = lightglue_ref.scatter_matches(m, ms, na, 0.0)
        assert c == cnt and np.array_equal(exp, vn)
    # the other overloads (SPmatcher.cc:359-410): all three normalise with 300 x 400; the Point2f + Mat one was handed a
    # NON-EMPTY vector (5 x 77): resize(n, -1) keeps those five entries unless a match overwrites them (SPmatcher.cc:375)
    m, ms = fe.match(ref[0][0], ref[1][0], ref[0][2], ref[1][2], 300, 400)
    exp, c = lightglue_ref.scatter_matches(m, ms, na, 0.0)
    assert m_pf == c and m_fp == c and m_kpinf == c and n_vn_pf == na
    stale = exp.copy()
    stale[:5] = np.where(exp[:5] >= 0, exp[:5], 77)
    assert np.array_equal(vn_pf, stale)
    assert np.array_equal(vn_kpinf, exp)                 # Matcher_Inference(KeyPoint...) fed normalised points, after an odd PreProcess call
    assert n_app_ret == na + 3 and n_app_ok == 1         # operator() appends and returns the vector's size (superpoint_onnx.cc:230)
    # NormalizeImage on BGR bytes (5, 25, 45): RGB order, 1/255; RGB2Grayscale with OpenCV's weights
    assert np.allclose(px0[:3], np.float32([45, 25, 5]) * np.float32(1 / 255.0), atol=1e-7)
    assert abs(px0[3] - float(np.float32([45, 25, 5]) @ np.float32([0.299, 0.587, 0.114])) / 255.0) < 1e-6
    rm, rms = lightglue_ref.LightGlueRef()(lightglue_ref.normalize_keypoints(ref[0][0], h, w),
                                           lightglue_ref.normalize_keypoints(ref[1][0], h, w), ref[0][2], ref[1][2])
    m, ms = fe.match(ref[0][0], ref[1][0], ref[0][2], ref[1][2], h, w)
    parity.compare_matches(rm.numpy(), rms.numpy(), m, ms)
    # 8(f).4: the adaptive score rule (superpoint_onnx.cc:192-210) switched on with lastmatch = 150: exactly the oracle's subset
    from oracle import frontend_aux_ref as aux
    keep = aux.adaptive_filter(ref[0][1], 150.0)
    assert n_ada == len(keep), (n_ada, len(keep), na, float(aux.adaptive_threshold(ref[0][1], 150.0)))
    assert np.array_equal(kp_ada[:, :2], ref[0][0][keep].astype(np.float32)) and np.array_equal(kp_ada[:, 2], ref[0][1][keep])
    assert np.array_equal(de_ada, ref[0][2][keep])
    fe.close()
