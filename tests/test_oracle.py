"""CPU tests: the oracle restatements against the committed golden vectors (which were produced by a
literal execution of the reference's ONNX graphs, tests/golden/make_golden.py), and -- when the
reference is mounted -- against that literal execution directly."""
import os

import numpy as np
import pytest
import torch

from oracle import lightglue_ref, superpoint_ref, synth
from tests import parity

REF_ONNX = "/root/reference/onnxmodel"


@pytest.fixture(scope="module")
def sp():
    return superpoint_ref.SuperPointRef()


@pytest.fixture(scope="module")
def lg():
    return lightglue_ref.LightGlueRef()


@pytest.mark.parametrize("name,seed,hw", [("sp_640x480_seed0", 0, (480, 640))])
def test_synth_reproduces_golden_image(golden_dir, name, seed, hw):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    assert np.array_equal(synth.frame(seed, *hw), g["image"])


@pytest.mark.parametrize("name", ["sp_640x480_seed0", "sp_752x480_seed100_a", "sp_752x480_seed100_b"])
def test_superpoint_restatement_vs_golden(golden_dir, sp, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    taps = {}
    k, s, d = sp(g["image"], taps)
    r = parity.compare_keypoints(g["keypoints"], g["scores"], k.numpy(), s.numpy())
    assert len(r["only_ref"]) == 0 and len(r["only_tst"]) == 0      # same build of torch: exact set
    heat = taps["heatmap"][0].numpy()
    assert np.abs(heat[g["heat_rows"]] - g["heat"]).max() <= parity.HEAT_ATOL
    parity.compare_descriptors(g["desc"], d.numpy()[g["desc_rows"]])
    dense = taps["dense_desc"][0].numpy()[:, ::8, ::8]
    assert np.abs(dense - g["dense_desc_px"]).max() <= parity.DESC_ATOL


def test_superpoint_properties(sp):
    img = synth.frame(3, 120, 160)
    k, s, d = sp(img)
    assert k.shape[0] > 10
    assert (k[:, 0] >= 4).all() and (k[:, 0] <= 160 - 5).all() and (k[:, 1] >= 4).all() and (k[:, 1] <= 120 - 5).all()
    assert (s > superpoint_ref.THRESHOLD).all()
    assert np.allclose(d.norm(dim=1).numpy(), 1.0, atol=1e-5)
    key = k[:, 1] * 100000 + k[:, 0]
    assert (key[1:] > key[:-1]).all()
    # NMS radius 4: no two keypoints within a 9x9 window of each other share... (Chebyshev distance > 4
    # is NOT guaranteed by the 2-iteration scheme, but distance 0 duplicates are impossible)
    assert len(set(map(tuple, k.tolist()))) == k.shape[0]


def test_superpoint_blank_image(sp):
    k, s, d = sp(np.zeros((64, 96), np.uint8))
    assert d.shape == (k.shape[0], 256)


@pytest.mark.parametrize("n", [256, 512])
def test_lightglue_restatement_vs_golden_synth(golden_dir, lg, n):
    g = np.load(os.path.join(golden_dir, f"lg_synth_n{n}.npz"))
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 200 + n)
    kn0 = lightglue_ref.normalize_keypoints(k0, 480, 640)
    kn1 = lightglue_ref.normalize_keypoints(k1, 480, 640)
    taps = {}
    m, ms = lg(kn0, kn1, d0, d1, taps)
    r = parity.compare_matches(g["matches"], g["mscores"], m.numpy(), ms.numpy())
    assert r["only_ref"] == 0 and r["only_tst"] == 0
    S = taps["S"].numpy()
    assert np.abs(S[0] - g["S_row0"]).max() < 2e-2 * max(1.0, np.abs(g["S_row0"]).max() / 50)
    # known answer: the second set is a permutation of the first
    inv = np.argsort(perm)
    assert (m[:, 1].numpy() == inv[m[:, 0].numpy()]).mean() > 0.99


def test_lightglue_pair_vs_golden(golden_dir, sp, lg):
    ga = np.load(os.path.join(golden_dir, "sp_752x480_seed100_a.npz"))
    gb = np.load(os.path.join(golden_dir, "sp_752x480_seed100_b.npz"))
    g = np.load(os.path.join(golden_dir, "lg_752x480_seed100.npz"))
    ka, sa, da = sp(ga["image"])
    kb, sb, db = sp(gb["image"])
    assert int(g["n0"]) == len(ka) and int(g["n1"]) == len(kb)
    kn0 = lightglue_ref.normalize_keypoints(ka.numpy(), 480, 752)
    kn1 = lightglue_ref.normalize_keypoints(kb.numpy(), 480, 752)
    m, ms = lg(kn0, kn1, da, db)
    parity.compare_matches(g["matches"], g["mscores"], m.numpy(), ms.numpy())
    disp = (kb[m[:, 1]] - ka[m[:, 0]]).numpy()
    assert tuple(np.median(disp, 0)) == (12.0, 7.0)


def test_lightglue_empty(lg):
    m, ms = lg(np.zeros((0, 2), np.float32), np.zeros((5, 2), np.float32),
               np.zeros((0, 256), np.float32), np.zeros((5, 256), np.float32))
    assert m.shape == (0, 2) and ms.shape == (0,)


def test_scatter_matches():
    vn, c = lightglue_ref.scatter_matches([[0, 3], [2, 1]], [0.5, 0.05], 4, 0.1)
    assert vn.tolist() == [3, -1, -1, -1] and c == 1
    vn, c = lightglue_ref.scatter_matches([[0, 3], [2, 1]], [0.5, 0.05], 4, 0.0)
    assert vn.tolist() == [3, -1, 1, -1] and c == 2


@pytest.mark.skipif(not os.path.isdir(REF_ONNX), reason="reference not mounted")
def test_restatement_vs_literal_onnx_execution(sp, lg):
    from oracle import onnx_interp
    spi = onnx_interp.Interpreter(os.path.join(REF_ONNX, "superpoint.onnx"))
    lgi = onnx_interp.Interpreter(os.path.join(REF_ONNX, "lightglue_sim.onnx"))
    a, b = synth.frame_pair(5, 240, 320, shift=(6, -3))
    feats = []
    for img in (a, b):
        x = torch.from_numpy(img.astype(np.float32) * np.float32(1 / 255.0))[None, None]
        o = spi.run({"image": x})
        k, s, d = sp(img)
        r = parity.compare_keypoints(o["keypoints"][0].numpy(), o["scores"][0].numpy(), k.numpy(), s.numpy(), strict=True)
        parity.compare_descriptors(o["descriptors"][0].numpy(), d.numpy())
        feats.append((k.numpy(), d.numpy()))
    kn0 = lightglue_ref.normalize_keypoints(feats[0][0], 240, 320)
    kn1 = lightglue_ref.normalize_keypoints(feats[1][0], 240, 320)
    o = lgi.run({"kpts0": kn0[None], "kpts1": kn1[None], "desc0": feats[0][1][None], "desc1": feats[1][1][None]})
    m, ms = lg(kn0, kn1, feats[0][1], feats[1][1])
    r = parity.compare_matches(o["matches0"].numpy(), o["mscores0"].numpy(), m.numpy(), ms.numpy())
    assert r["only_ref"] == 0 and r["only_tst"] == 0
