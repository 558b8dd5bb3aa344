"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; argument checking fails loudly; there is no CPU fallback."""
import ctypes
import os
import subprocess

import pytest

from rover_slam_b200 import api

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.lib_path()):
        subprocess.run(["make", "-j8", "-C", ROOT], check=True)
    return api.load_library()


def test_header_symbols_exported(lib):
    syms = api.exported_symbols()
    assert {"rfe_create", "rfe_destroy", "rfe_sp_extract_u8", "rfe_lg_match", "rfe_last_error"} <= set(syms)
    for s in syms:
        assert getattr(lib, s) is not None


def test_library_has_blackwell_tensor_core_code():
    out = subprocess.run(["cuobjdump", "-sass", api.lib_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in out, "TMA loads missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"
    assert "sm_100a" in out


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.RoverFeError) as e:
        api.FrontEnd()
    assert e.value.code == api.RFE_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_null_arguments(lib):
    assert lib.rfe_create(None, None) == api.RFE_ERR_INVALID
    assert b"null" in lib.rfe_last_error()
    assert lib.rfe_sync(None) == api.RFE_ERR_INVALID
    lib.rfe_destroy(None)   # no-op


def test_product_does_not_import_oracle():
    """The product path (package + csrc) must never reference oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rover_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the CPU oracle", ""), f"{f} mentions oracle"


def test_weight_blob_matches_reference_when_mounted():
    """EVERY tensor of weights/rover_fe.rfw against the initialisers of the reference's two ONNX files:
    (1) layout-independent: the multiset of values of each blob tensor (hash of the sorted values) is the multiset of exactly one
        ONNX float initialiser and vice versa -- nothing missing, nothing invented, nothing altered;
    (2) layout: every tensor equals what the documented repacking (OIHW -> OHWI, [in,out] -> [out,in]) yields."""
    ref = "/root/reference/onnxmodel"
    if not os.path.exists(ref + "/superpoint.onnx"):
        pytest.skip("reference not mounted")
    import hashlib
    import sys
    import numpy as np
    from oracle import onnx_reader, weights
    blob = weights.load()

    def h(a):
        return hashlib.sha1(np.sort(np.asarray(a, np.float32).reshape(-1)).tobytes()).hexdigest()

    blob_h = {h(v): k for k, v in blob.items()}
    onnx_h, n_params = {}, 0
    for f in ("superpoint.onnx", "lightglue_sim.onnx"):
        g = onnx_reader.load(os.path.join(ref, f))
        for k, v in g.initializers.items():
            if v.dtype == np.float32 and v.size >= 2:
                onnx_h[h(v)] = f + ":" + k
                n_params += v.size
    assert len(blob) == 227 and len(blob_h) == 227
    assert not [v for k, v in onnx_h.items() if k not in blob_h]                       # every reference weight is in the blob
    extra = [v for k, v in blob_h.items() if k not in onnx_h]
    assert extra == ["lg.matchability.b"]                                              # the one scalar initialiser (size 1)
    g2 = onnx_reader.load(os.path.join(ref, "lightglue_sim.onnx"))
    assert np.array_equal(blob["lg.matchability.b"].reshape(-1), g2.initializers["log_assignment.8.matchability.bias"].reshape(-1))
    assert sum(v.size for v in blob.values()) == n_params + 1
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pack_weights
    want = {}
    want.update(pack_weights.superpoint_tensors(os.path.join(ref, "superpoint.onnx")))
    want.update(pack_weights.lightglue_tensors(os.path.join(ref, "lightglue_sim.onnx")))
    assert sorted(want) == sorted(blob)
    for k, v in want.items():
        assert np.array_equal(blob[k], np.asarray(v, np.float32)), k
    g = onnx_reader.load(os.path.join(ref, "superpoint.onnx"))
    assert np.array_equal(blob["sp.conv3b.w"], g.initializers["conv3b.weight"].transpose(0, 2, 3, 1))
    assert np.array_equal(blob["lg.l4.cross.to_v.b"], g2.initializers["transformers.4.cross_attn.to_v.bias"])
    assert blob["lg.l0.self.wqkv.w"].shape == (768, 256)


def test_host_classes_without_device_follow_the_reference_error_convention(tmp_path):
    """errors never cross the class surface as exceptions (the reference returns EXIT_FAILURE,
    superpoint_onnx.cc:62-65,158-161): without a GPU SPextractor::operator() yields 0 keypoints and
    SPmatcher::MatchingPoints_onnx 0 matches, each failure is reported on stderr, and nothing falls back to the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    drv = os.path.join(ROOT, "rover_slam_b200", "host_driver")
    if not os.path.exists(drv):
        subprocess.run(["make", "-C", ROOT, "host"], check=True, capture_output=True)
    a, b, out = (str(tmp_path / n) for n in ("a.raw", "b.raw", "out.bin"))
    img = (np.arange(240 * 320) % 251).astype(np.uint8)
    img.tofile(a)
    img.tofile(b)
    r = subprocess.run([drv, "240", "320", a, b, out], capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, ROVER_FE_WEIGHTS=os.path.join(ROOT, "weights", "rover_fe.rfw")))
    assert r.returncode == 0, r.stderr
    assert "no CPU fallback" in r.stderr and "init failed" in r.stderr
    hdr = np.fromfile(out, dtype=np.int32)[:6]
    assert hdr.tolist() == [0, 0, 0, 0, 0, 0]            # keypoints a / b, multi-level, matches (Frame / KeyPoint overload), adaptive
