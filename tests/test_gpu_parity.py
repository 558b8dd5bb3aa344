"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
and the committed golden vectors.  Tolerances: tests/parity.py."""
import os

import numpy as np
import pytest
import torch

from oracle import lightglue_ref, superpoint_ref, synth
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    from rover_slam_b200 import FrontEnd
    f = FrontEnd(max_batch=8, max_height=480, max_width=752, max_keypoints=4096)
    yield f
    f.close()


@pytest.fixture(scope="module")
def sp():
    return superpoint_ref.SuperPointRef()


@pytest.fixture(scope="module")
def lg():
    return lightglue_ref.LightGlueRef()


# ---- the tensor-core building block -------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (128, 128, 512), (300, 200, 512), (77, 768, 256), (1, 8, 8),
                                   (2049, 65, 256)])
def test_split_fp16_gemm_is_fp32_equivalent(fe, m, n, k):
    rng = np.random.RandomState(m + n + k)
    a = (rng.randn(m, k) * 3).astype(np.float32)
    b = rng.randn(n, k).astype(np.float32)
    bias = rng.randn(n).astype(np.float32)
    d = fe.debug_gemm(a, b, bias)
    ref64 = a.astype(np.float64) @ b.astype(np.float64).T + bias
    ref32 = a @ b.T + bias
    err = np.abs(d - ref64).max()
    err32 = np.abs(ref32 - ref64).max()
    scale = np.abs(ref64).max()
    assert err <= max(4 * err32, 2e-6 * scale), (err, err32)


# ---- SuperPoint -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sp_640x480_seed0", "sp_752x480_seed100_a"])
def test_superpoint_vs_golden(fe, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    (k, s, d), = fe.extract(g["image"])
    r = parity.compare_keypoints(g["keypoints"], g["scores"], k, s, strict=True)      # identical sets on the goldens
    h, w = g["image"].shape
    heat = fe.debug_read("sp.heat").reshape(-1, h, w)[0]
    assert np.abs(heat[g["heat_rows"]] - g["heat"]).max() <= parity.HEAT_ATOL
    # descriptors of the golden subset that is common
    pos = {int(i): j for j, i in enumerate(r["ref_idx"])}
    rows = [(gi, r["tst_idx"][pos[int(ri)]]) for gi, ri in enumerate(g["desc_rows"]) if int(ri) in pos]
    gi, ti = np.array(rows).T
    parity.compare_descriptors(g["desc"][gi], d[ti])
    dense = fe.debug_read("sp.dense").reshape(h // 8, w // 8, 256)
    assert np.abs(np.transpose(dense, (2, 0, 1))[:, ::8, ::8] - g["dense_desc_px"]).max() <= parity.DESC_ATOL


def test_superpoint_batch8_vs_oracle(fe, sp):
    """BASELINE config 2: batch of 8 synthetic 640x480 frames."""
    imgs = np.stack([synth.frame(s, 480, 640) for s in range(8)])
    feats = fe.extract(imgs)
    heat = fe.debug_read("sp.heat").reshape(8, 480, 640)
    for i in range(8):
        rk, rs, rd = sp(imgs[i])
        k, s, d = feats[i]
        r = parity.compare_keypoints(rk.numpy(), rs.numpy(), k, s, heat=heat[i])
        parity.compare_descriptors(rd.numpy()[r["ref_idx"]], d[r["tst_idx"]])
        assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)


def test_superpoint_intermediates(fe, sp):
    img = synth.frame(21, 96, 160)
    fe.extract(img)
    taps = {}
    sp(img, taps)
    for name, key, c, div, tol in [("sp.a1a", "relu1a", 64, 1, 2e-4), ("sp.pool1", "pool1", 64, 2, 2e-4),
                                   ("sp.pool2", "pool2", 64, 4, 2e-4), ("sp.pool3", "pool3", 128, 8, 2e-4),
                                   ("sp.feat", "feat", 128, 8, 2e-4)]:
        got = np.transpose(fe.debug_read(name).reshape(96 // div, 160 // div, c), (2, 0, 1))
        ref = taps[key][0].numpy()
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), name


def test_nms_kernel_exact_on_own_heatmap(fe):
    """NMS + border + threshold + ordering are integer/comparison work: bit-exact given the same heat-map."""
    img = synth.frame(5, 240, 320)
    (k, s, _), = fe.extract(img, want_desc=False)
    heat = torch.from_numpy(fe.debug_read("sp.heat").reshape(1, 240, 320))
    nmsed = superpoint_ref.SuperPointRef.nms(heat)
    rk, rs, post = superpoint_ref.SuperPointRef.select(nmsed)
    # the map itself: identical wherever a keypoint can come from (scores above the threshold) and on the -1 border; at
    # sub-threshold maxima the graph's map holds the score and ours 0 -- the graph's own `s > 0.0005` selection ignores both
    got, want = fe.debug_read("sp.nms").reshape(240, 320), post[0].numpy()
    sel = (want > superpoint_ref.THRESHOLD) | (got > superpoint_ref.THRESHOLD) | (want < 0) | (got < 0)
    assert np.array_equal(got[sel], want[sel]) and sel.sum() > 2000
    assert (got[~sel] == 0).all()
    assert np.array_equal(rk.numpy(), k)
    assert np.array_equal(rs.numpy(), s)


@pytest.mark.parametrize("hw", [(8, 8), (16, 24), (64, 96), (480, 752)])
def test_superpoint_sizes_and_blank(fe, sp, hw):
    h, w = hw
    for img in (np.zeros((h, w), np.uint8), synth.frame(9, h, w) if min(h, w) >= 16 else np.full((h, w), 200, np.uint8)):
        (k, s, d), = fe.extract(img)
        rk, rs, rd = sp(img)
        r = parity.compare_keypoints(rk.numpy(), rs.numpy(), k, s, heat=fe.debug_read("sp.heat").reshape(-1, h, w)[0])
        if len(r["ref_idx"]):
            parity.compare_descriptors(rd.numpy()[r["ref_idx"]], d[r["tst_idx"]])


def test_superpoint_rejects_bad_sizes(fe):
    from rover_slam_b200 import RoverFeError
    with pytest.raises(RoverFeError):
        fe.extract(np.zeros((100, 100), np.uint8))       # not multiples of 8
    with pytest.raises(RoverFeError):
        fe.extract(np.zeros((9, 480, 640), np.uint8))    # batch > max_batch


# ---- LightGlue --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [256, 512])
def test_lightglue_vs_golden_synth(fe, golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"lg_synth_n{n}.npz"))
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 200 + n)
    m, ms = fe.match(k0, k1, d0, d1, 480, 640)
    r = parity.compare_matches(g["matches"], g["mscores"], m, ms)
    assert r["only_ref"] == 0 and r["only_tst"] == 0          # identical pairs on the goldens


@pytest.mark.parametrize("n0,n1", [(1024, 1024), (300, 777), (5, 1000), (1, 1), (2048, 2048)])
def test_lightglue_vs_oracle_ragged(fe, lg, n0, n1):
    """BASELINE config 4 (N sweep) + ragged sizes."""
    n = max(n0, n1)
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 300 + n)
    k0, d0, k1, d1 = k0[:n0], d0[:n0], k1[:n1], d1[:n1]
    m, ms = fe.match(k0, k1, d0, d1, 480, 640)
    rm, rms = lg(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1)
    parity.compare_matches(rm.numpy(), rms.numpy(), m, ms)


_BRES_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import synth
from rover_slam_b200 import FrontEnd
fe = FrontEnd(max_batch=2, max_height=64, max_width=64, max_keypoints=2048)
out = {}
for n0, n1 in ((1024, 1024), (300, 777), (1, 1)):
    n = max(n0, n1)
    k0, k1, d0, d1, _ = synth.lightglue_inputs(n, 300 + n)
    m, ms = fe.match(k0[:n0], k1[:n1], d0[:n0], d1[:n1], 480, 640)
    out[f"m_{n0}_{n1}"], out[f"s_{n0}_{n1}"] = m, ms
np.savez(sys.argv[2], **out)
"""


@pytest.mark.parametrize("mode", ["0", "2"])
def test_lightglue_gemm_tile_modes_vs_oracle(lg, tmp_path, mode):
    """The 256 -> 256 linears and Wqkv have two tilings (umma_kernel.cuh): streamed 64-wide tiles below one wave of tiles,
    128-wide B-resident tiles above.  RFE_BRES=0 / 2 (read when the library loads, hence the child process) forces either
    one for every size; both must agree with the oracle."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "bres.npz")
    env = dict(os.environ, RFE_BRES=mode)
    subprocess.run([sys.executable, "-c", _BRES_CHILD, root, out], check=True, env=env, timeout=600)
    got = np.load(out)
    for n0, n1 in ((1024, 1024), (300, 777), (1, 1)):
        n = max(n0, n1)
        k0, k1, d0, d1, _ = synth.lightglue_inputs(n, 300 + n)
        k0, d0, k1, d1 = k0[:n0], d0[:n0], k1[:n1], d1[:n1]
        rm, rms = lg(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1)
        parity.compare_matches(rm.numpy(), rms.numpy(), got[f"m_{n0}_{n1}"], got[f"s_{n0}_{n1}"])


def test_bench_shape_batch_vs_oracle(sp, lg):
    """BASELINE config 5 shape: 8 pairs of 640x480 frames in ONE batched pass (16 frames extracted, 8 pairs matched: the
    size at which the B-resident GEMM tiles and the persistent attention launches are used).  ALL 8 pairs against the
    oracle: the 16 keypoint sets, then the matches of the oracle run on OUR features (so a keypoint near-tie cannot
    excuse a matcher difference)."""
    from rover_slam_b200 import FrontEnd
    import bench
    frames = bench.make_pairs(8, 3)
    big = FrontEnd(max_batch=16, max_height=480, max_width=640, max_keypoints=4096)
    try:
        kpts, res = big.match_pairs(frames.reshape(16, 480, 640))
        kpts = [k.copy() for k in kpts]
        res = [(m.copy(), s.copy()) for m, s in res]
        heat = big.debug_read("sp.heat").reshape(16, 480, 640)
        feats = [big.read_slot(b) for b in range(16)]          # the features the batched matcher consumed
    finally:
        big.close()
    for pi in range(8):
        for f in (0, 1):
            rk, rs, rd = sp(frames[pi, f])
            k, s, d = feats[2 * pi + f]
            assert np.array_equal(k, kpts[2 * pi + f])
            r = parity.compare_keypoints(rk.numpy(), rs.numpy(), k, s, heat=heat[2 * pi + f])
            parity.compare_descriptors(rd.numpy()[r["ref_idx"]], d[r["tst_idx"]])
        (k0, _, d0), (k1, _, d1) = feats[2 * pi], feats[2 * pi + 1]
        rm, rms = lg(lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640), d0, d1)
        parity.compare_matches(rm.numpy(), rms.numpy(), res[pi][0], res[pi][1])
        assert len(res[pi][0]) > 500


def test_lightglue_empty_inputs(fe):
    k0, k1, d0, d1, _ = synth.lightglue_inputs(16, 1)
    m, ms = fe.match(k0[:0], k1, d0[:0], d1, 480, 640)
    assert m.shape == (0, 2) and ms.shape == (0,)
    m, ms = fe.match(k0, k1[:0], d0, d1[:0], 480, 640)
    assert m.shape == (0, 2)


def test_lightglue_threshold_and_normalisation(fe, lg):
    k0, k1, d0, d1, _ = synth.lightglue_inputs(256, 9)
    m_all, s_all = fe.match(k0, k1, d0, d1, 480, 640, thresh=0.0)
    m_hi, s_hi = fe.match(k0, k1, d0, d1, 480, 640, thresh=0.5)
    assert (s_hi > 0.5).all() and len(m_hi) <= len(m_all)
    keep = s_all > 0.5
    assert np.array_equal(m_all[keep], m_hi)
    # the three non-Frame overloads of the reference hard-code rows=300, cols=400 (SPmatcher.cc:360-361)
    m3, s3 = fe.match(k0, k1, d0, d1, 300, 400)
    rm, rms = lg(lightglue_ref.normalize_keypoints(k0, 300, 400), lightglue_ref.normalize_keypoints(k1, 300, 400), d0, d1)
    parity.compare_matches(rm.numpy(), rms.numpy(), m3, s3)


# ---- end to end (BASELINE config 3) ------------------------------------------------------------------------
def test_pair_end_to_end_752x480(fe, golden_dir, sp, lg):
    ga = np.load(os.path.join(golden_dir, "sp_752x480_seed100_a.npz"))
    gb = np.load(os.path.join(golden_dir, "sp_752x480_seed100_b.npz"))
    g = np.load(os.path.join(golden_dir, "lg_752x480_seed100.npz"))
    imgs = np.stack([ga["image"], gb["image"]])
    feats = fe.extract(imgs)
    ra = parity.compare_keypoints(ga["keypoints"], ga["scores"], feats[0][0], feats[0][1])
    rb = parity.compare_keypoints(gb["keypoints"], gb["scores"], feats[1][0], feats[1][1])
    m, ms = fe.match(feats[0][0], feats[1][0], feats[0][2], feats[1][2], 480, 752)
    if len(ra["only_ref"]) + len(ra["only_tst"]) + len(rb["only_ref"]) + len(rb["only_tst"]) == 0:
        parity.compare_matches(g["matches"], g["mscores"], m, ms)      # identical features -> golden matches
    # in every case: identical to the oracle run on OUR features
    rm, rms = lg(lightglue_ref.normalize_keypoints(feats[0][0], 480, 752),
                 lightglue_ref.normalize_keypoints(feats[1][0], 480, 752), feats[0][2], feats[1][2])
    parity.compare_matches(rm.numpy(), rms.numpy(), m, ms)
    disp = feats[1][0][m[:, 1]] - feats[0][0][m[:, 0]]
    assert tuple(np.median(disp, 0)) == (12.0, 7.0)
    # device-resident hand-off gives the same answer as the host round trip
    fe.match_slots(0, 1, 480, 752)
    m2, ms2 = fe.read_result(0)
    assert np.array_equal(m, m2) and np.array_equal(ms, ms2)


def test_batched_matching_equals_single(fe):
    """rfe_lg_match_slots_batch (all pairs in one pass) == one rfe_lg_match_slots per pair, bit for bit."""
    imgs = np.stack([f for s in (40, 41, 42) for f in synth.frame_pair(s, 240, 320, shift=(5 + s % 3, -2))])
    h, w = 240, 320
    fe.extract_device_from_host(imgs)
    single = []
    for p in range(3):
        fe.match_slots(2 * p, 2 * p + 1, h, w, 0.0, p)
        single.append(fe.read_result(p))
    fe.match_slots_batch([0, 2, 4], [1, 3, 5], h, w)
    for p in range(3):
        m, ms = fe.read_result(p)
        assert np.array_equal(m, single[p][0]) and np.array_equal(ms, single[p][1])
        assert len(m) > 50


def test_match_pairs_one_call_pipeline(fe):
    """rfe_match_pairs_u8 (host frames in, keypoints + matches out) == extract + match through the per-call API."""
    imgs = np.stack([f for s in (50, 51) for f in synth.frame_pair(s, 240, 320, shift=(4, 6))])
    kpts, res = fe.match_pairs(imgs)
    kpts = [k.copy() for k in kpts]
    res = [(m.copy(), s.copy()) for m, s in res]
    feats = fe.extract(imgs)
    for p in range(2):
        assert np.array_equal(kpts[2 * p], feats[2 * p][0]) and np.array_equal(kpts[2 * p + 1], feats[2 * p + 1][0])
        m, ms = fe.match(feats[2 * p][0], feats[2 * p + 1][0], feats[2 * p][2], feats[2 * p + 1][2], 240, 320)
        assert np.array_equal(res[p][0], m) and np.array_equal(res[p][1], ms)


def test_write_slot_round_trip_and_match(fe):
    """rfe_sp_write_slot (8(f).3: a stored KeyFrame re-enters the device-resident path): the slot reads back bit for bit,
    its binarised copy follows, and matching uploaded slots == rfe_lg_match on the same host features, bit for bit --
    also against a slot that a real extraction filled."""
    from rover_slam_b200.api import RoverFeError
    a, b = synth.frame_pair(61, 240, 320, shift=(6, 3))
    fa, fb = fe.extract(np.stack([a, b]))
    (ka, sa, da), (kb, sb, db) = [(x[0].copy(), x[1].copy(), x[2].copy()) for x in (fa, fb)]
    m_ref, s_ref = fe.match(ka, kb, da, db, 240, 320)
    assert len(m_ref) > 50
    fe.extract_device_from_host(np.stack([a, b]))        # slots 0, 1 = the same two frames
    fe.write_slot(3, kb, db, sb)                          # slot 2 is skipped: it must read as empty
    k3, s3, d3 = fe.read_slot(3)
    assert np.array_equal(k3, kb) and np.array_equal(s3, sb) and np.array_equal(d3, db)
    assert np.array_equal(fe.read_slot_bin(3), (db > 0).astype(np.uint8))
    assert len(fe.read_slot(2)[0]) == 0
    fe.match_slots(0, 3, 240, 320, 0.0, 0)                # extracted slot vs uploaded slot
    m, ms = fe.read_result(0)
    assert np.array_equal(m, m_ref) and np.array_equal(ms, s_ref)
    fe.write_slot(0, ka, da)                              # scores are optional
    fe.match_slots_batch([0, 3], [3, 0], 240, 320)
    m, ms = fe.read_result(0)
    assert np.array_equal(m, m_ref) and np.array_equal(ms, s_ref)
    m10, _ = fe.match(kb, ka, db, da, 240, 320)
    assert np.array_equal(fe.read_result(1)[0], m10)
    fe.write_slot(1, ka[:0], da[:0])                      # empty upload
    assert len(fe.read_slot(1)[0]) == 0
    with pytest.raises(RoverFeError):
        fe.write_slot(8, ka, da)                          # slot >= max_batch
    with pytest.raises(RoverFeError):
        fe.write_slot(0, np.zeros((5000, 2)), np.zeros((5000, 256), np.float32))   # n > max_keypoints


def test_concurrent_contexts_from_threads(fe):
    """SURVEY 8(b) threading contract: one ctx per SLAM thread (Tracking, LocalMapping, LoopClosing each own an extractor and a
    matcher), used concurrently.  Three threads with their own ctx (own stream, buffers, thread-local tensor-map cache and error
    string) must return exactly what a single thread returns for the same frames, every repetition."""
    import threading
    from rover_slam_b200 import FrontEnd
    pairs = [np.stack(synth.frame_pair(70 + i, 240, 320, shift=(3 + i, 2 - i))) for i in range(3)]
    want = []
    for imgs in pairs:
        kp, res = fe.match_pairs(imgs)
        want.append(([k.copy() for k in kp], res[0][0].copy(), res[0][1].copy()))
    errors = []

    def worker(i):
        try:
            ctx = FrontEnd(max_batch=2, max_height=240, max_width=320, max_keypoints=2048)
            for _ in range(4):
                kp, res = ctx.match_pairs(pairs[i])
                assert all(np.array_equal(a, b) for a, b in zip(kp, want[i][0]))
                assert np.array_equal(res[0][0], want[i][1]) and np.array_equal(res[0][1], want[i][2])
            ctx.close()
        except Exception as e:          # noqa: BLE001 -- reported by the main thread
            errors.append((i, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors
    assert len(want[0][1]) > 50


def test_gpu_path_launches_kernels(fe):
    before = fe.kernel_launches()
    fe.extract(synth.frame(1, 64, 64))
    assert fe.kernel_launches() - before >= 15


def test_pipelined_pairs_equal_one_shot_calls(fe):
    """rfe_pairs_submit / rfe_pairs_collect with two batches in flight return exactly what rfe_match_pairs_u8 returns."""
    batches = [np.stack([f for s in range(2) for f in synth.frame_pair(40 + 10 * b + s, 240, 320, shift=(4, 2))]) for b in range(3)]
    ref = []
    for imgs in batches:
        kp, res = fe.match_pairs(imgs)
        ref.append(([k.copy() for k in kp], [(m.copy(), s.copy()) for m, s in res]))
    out = []
    fe.pairs_submit(batches[0])
    for i in range(3):
        if i + 1 < 3:
            fe.pairs_submit(batches[i + 1])
        kp, res = fe.pairs_collect()
        out.append(([k.copy() for k in kp], [(m.copy(), s.copy()) for m, s in res]))
    # ... and the begin / submit / end ordering that keeps the GPU queue full
    out2 = []
    fe.pairs_submit(batches[0])
    fe.pairs_submit(batches[1])
    for i in range(3):
        fe.pairs_collect_begin()
        if i + 2 < 3:
            fe.pairs_submit(batches[i + 2])
        kp, res = fe.pairs_collect_end()
        out2.append(([k.copy() for k in kp], [(m.copy(), s.copy()) for m, s in res]))
    for (rk, rr), (ok, orr) in zip(ref + ref, out + out2):
        for a, b in zip(rk, ok):
            assert np.array_equal(a, b)
        for (rm, rs), (om, os_) in zip(rr, orr):
            assert np.array_equal(rm, om) and np.array_equal(rs, os_)
            assert len(rm) > 20



def test_padding_rows_do_not_drift_over_many_calls(fe):
    """Images start at multiples of 8 rows; the padding rows go through every GEMM / FFN of every call.  They are reset by
    lg_prepare on each call: 3000 matches with n % 8 != 0 return exactly what the first one returned, and the residual
    stream of a padding row stays bounded (it would otherwise grow by ~11 per call until the fp16 planes overflow)."""
    k0, k1, d0, d1, _ = synth.lightglue_inputs(64, 77)
    k0, d0, k1, d1 = k0[:13], d0[:13], k1[:27], d1[:27]
    first = fe.match(k0, k1, d0, d1, 480, 640)
    assert len(first[0]) >= 2
    for _ in range(3000):
        m, ms = fe.match(k0, k1, d0, d1, 480, 640)
    assert np.array_equal(m, first[0]) and np.array_equal(ms, first[1])
    x = fe.debug_read("lg.x").reshape(-1, 256)            # rows [0, 13) image 0, [13, 16) padding, [16, 43) image 1
    assert np.isfinite(x).all() and np.abs(x[13:16]).max() < 1e3


def test_match_normalized_equals_match(fe):
    """rfe_lg_match_normalized (the host runner's path: keypoints normalised on the host exactly like
    NormalizeKeypoints, transform.cpp:19-32) == rfe_lg_match (pixels, normalised on the device)."""
    k0, k1, d0, d1, _ = synth.lightglue_inputs(300, 5)
    for nh, nw in ((480, 640), (300, 400)):
        m, ms = fe.match(k0, k1, d0, d1, nh, nw)
        m2, ms2 = fe.match_normalized(lightglue_ref.normalize_keypoints(k0, nh, nw), lightglue_ref.normalize_keypoints(k1, nh, nw), d0, d1)
        assert np.array_equal(m, m2) and np.array_equal(ms, ms2)


def test_topk_cap_keeps_the_best_in_row_major_order(fe):
    """8(f).4: rfe_sp_set_topk(k) == the oracle's order-preserving top-k of the uncapped extraction, bit for bit (keypoints,
    scores, descriptors), including exact score ties at the cut; k >= N and k <= 0 leave everything."""
    from oracle import frontend_aux_ref as aux
    imgs = np.stack([synth.frame(s, 240, 320) for s in (8, 9)])
    full = fe.extract(imgs)
    try:
        for k in (1, 100, 257, len(full[0][0]) - 1, len(full[0][0]), 4000):
            fe.set_topk(k)
            got = fe.extract(imgs)
            for (fk, fs, fd), (gk, gs, gd) in zip(full, got):
                keep = aux.topk_keep_order(fs, k)
                assert np.array_equal(gk, fk[keep]) and np.array_equal(gs, fs[keep]) and np.array_equal(gd, fd[keep]), k
        # exact ties at the cut: a flat image region gives many identical scores? use the scores themselves: pick k so that
        # the k-th and (k+1)-th best scores are equal when such a pair exists
        fs = full[0][1]
        order = np.sort(fs)[::-1]
        ties = np.nonzero(order[:-1] == order[1:])[0]
        if len(ties):
            k = int(ties[0]) + 1
            fe.set_topk(k)
            gk, gs, gd = fe.extract(imgs[:1])[0]
            keep = aux.topk_keep_order(fs, k)
            assert np.array_equal(gk, full[0][0][keep]) and np.array_equal(gs, fs[keep])
    finally:
        fe.set_topk(0)
    again = fe.extract(imgs)
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(full, again))


_FUSE_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import synth
from rover_slam_b200 import FrontEnd
fe = FrontEnd(max_batch=2, max_height=240, max_width=328, max_keypoints=4096)
out = {}
for i, (h, w) in enumerate(((240, 328), (16, 24), (96, 136))):
    imgs = np.stack([synth.frame(60 + i, h, w), np.zeros((h, w), np.uint8)])
    for b, (k, s, d) in enumerate(fe.extract(imgs)):
        out[f"k{i}{b}"], out[f"s{i}{b}"], out[f"d{i}{b}"] = k, s, d
    out[f"p{i}"] = fe.debug_read("sp.pool1")
np.savez(sys.argv[2], **out)
"""


def test_fused_conv1a_is_bit_identical_to_the_two_kernel_path(tmp_path):
    """K1 fusion: conv1a computed inside conv1b's row producers (no activation round trip) == stand-alone conv1a kernel +
    conv1b, bit for bit (same fp32 FMA order, same split), on widths that are not multiples of 128 and on tiny images."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("0", "1"):
        out = str(tmp_path / f"fuse{mode}.npz")
        subprocess.run([sys.executable, "-c", _FUSE_CHILD, root, out], check=True, env=dict(os.environ, RFE_FUSE_CONV1A=mode), timeout=600)
        res[mode] = np.load(out)
    assert sorted(res["0"].files) == sorted(res["1"].files)
    for k in res["0"].files:
        assert np.array_equal(res["0"][k], res["1"][k]), k
    assert len(res["1"]["k00"]) > 100


def test_fast_mode_is_labelled_close_and_reversible(fe, sp):
    """rfe_set_fast_mode (hi-only fp16 MMAs in the backbone convolutions) is NOT the parity path: it must stay close to the
    exact path -- heat-map within 2e-3, at least 97 % of the keypoints in common, their descriptors within 2e-2 -- and
    switching it off must give back the exact result bit for bit."""
    img = synth.frame(5, 480, 640)
    exact = fe.extract(img)[0]
    heat_exact = fe.debug_read("sp.heat").copy()
    fe.set_fast_mode(True)
    try:
        fast = fe.extract(img)[0]
        heat_fast = fe.debug_read("sp.heat").copy()
    finally:
        fe.set_fast_mode(False)
    again = fe.extract(img)[0]
    assert all(np.array_equal(a, b) for a, b in zip(exact, again))
    assert 0 < np.abs(heat_fast - heat_exact).max() <= 2e-3          # it really is another arithmetic, and a close one
    ke = {(int(x), int(y)): i for i, (x, y) in enumerate(exact[0])}
    common = [(ke[(int(x), int(y))], j) for j, (x, y) in enumerate(fast[0]) if (int(x), int(y)) in ke]
    assert len(common) >= 0.97 * len(exact[0]) and len(common) >= 0.97 * len(fast[0])
    ie, jf = np.array(common).T
    assert np.abs(exact[2][ie] - fast[2][jf]).max() <= 2e-2


def test_sm_limit_changes_nothing_but_the_grid(fe):
    """rfe_set_sm_limit: the persistent kernels on 100 SMs give the same keypoints, descriptors and matches as on all of them
    (tile -> CTA assignment changes, per-tile arithmetic does not)."""
    a, b = synth.frame_pair(77, 480, 640)
    pair = np.stack([a, b])
    want_k, want_m = fe.match_pairs(pair)
    fe.set_sm_limit(100)
    try:
        got_k, got_m = fe.match_pairs(pair)
    finally:
        fe.set_sm_limit(0)
    assert all(np.array_equal(x, y) for x, y in zip(want_k, got_k))
    rep = parity.compare_matches(want_m[0][0], want_m[0][1], got_m[0][0], got_m[0][1])
    assert rep["only_ref"] + rep["only_tst"] <= 1 and rep["mscore_maxabs"] <= 1e-4, rep
    with pytest.raises(Exception):
        fe.set_sm_limit(100000)


_SPLIT_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from oracle import synth
from rover_slam_b200 import FrontEnd
fe = FrontEnd(max_batch=6, max_height=480, max_width=640, max_keypoints=1280)
out = {}
for tag, n, pairs in (("a", 1250, 2), ("b", 800, 3), ("c", 1250, 2)):     # c == a again: after a launch with another S
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 900 + n)
    for p in range(pairs):
        fe.write_slot(2 * p, k0, d0)
        fe.write_slot(2 * p + 1, k1, d1)
    fe.match_slots_batch(list(range(0, 2 * pairs, 2)), list(range(1, 2 * pairs, 2)), 480, 640)
    for p in range(pairs):
        out[f"m{tag}{p}"], out[f"s{tag}{p}"] = fe.read_result(p)
    out["perm" + tag] = perm
np.savez(sys.argv[2], **out)
"""


def test_attention_key_split_of_the_tail_items_equals_unsplit(tmp_path):
    """attn2_kernel cuts the work items of its last, partly filled round into key-range parts, merged by whichever part
    finishes last (2 pairs x 1250 keypoints: 160 items on 148 SMs -> 12 items in 4 parts; 3 pairs x 800: 168 items ->
    20 items in 3 parts).  Same matches as with RFE_ATTN_SPLIT=0, scores within the LightGlue tolerance, and every pair of the
    batch identical to the first (same inputs, different positions in the item list)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("0", "1"):
        out = str(tmp_path / f"split{mode}.npz")
        subprocess.run([sys.executable, "-c", _SPLIT_CHILD, root, out], check=True, env=dict(os.environ, RFE_ATTN_SPLIT=mode), timeout=600)
        res[mode] = np.load(out)
    for p in range(2):                              # bit-identical whichever part arrived last, and after S changed in between
        assert np.array_equal(res["1"][f"ma{p}"], res["1"][f"mc{p}"]) and np.array_equal(res["1"][f"sa{p}"], res["1"][f"sc{p}"])
    for tag, pairs in (("a", 2), ("b", 3)):
        for p in range(pairs):
            m0, s0 = res["0"][f"m{tag}{p}"], res["0"][f"s{tag}{p}"]
            m1, s1 = res["1"][f"m{tag}{p}"], res["1"][f"s{tag}{p}"]
            assert len(m0) > 500
            rep = parity.compare_matches(m0, s0, m1, s1)
            assert rep["only_ref"] + rep["only_tst"] <= 1 and rep["mscore_maxabs"] <= 1e-4, rep
        for p in range(1, pairs):               # same inputs at other positions of the item list (split or not)
            rep = parity.compare_matches(res["1"][f"m{tag}0"], res["1"][f"s{tag}0"], res["1"][f"m{tag}{p}"], res["1"][f"s{tag}{p}"])
            assert rep["only_ref"] + rep["only_tst"] <= 1 and rep["mscore_maxabs"] <= 1e-4, rep
        perm = res["1"]["perm" + tag]
        m1 = res["1"][f"m{tag}0"]
        assert (perm[m1[:, 1]] == m1[:, 0]).mean() > 0.99


# ---- against an independent runtime: OpenCV DNN executing the reference's ONNX files (tests/golden/cv2dnn_*.npz) ----------
@pytest.mark.parametrize("name", ["sp_640x480_seed0", "sp_752x480_seed100_a"])
def test_superpoint_vs_cv2dnn_golden(fe, golden_dir, name):
    """Heat-map (softmax-65 + depth-to-space) and L2-normalised dense descriptors of the CUDA path against the tensors
    cv2.dnn computed from superpoint.onnx (no code of this repository between the reference's graph and the fixture)."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    c = np.load(os.path.join(golden_dir, "cv2dnn_" + name + ".npz"))
    fe.extract(g["image"])
    h, w = g["image"].shape
    heat = fe.debug_read("sp.heat").reshape(-1, h, w)[0]
    assert np.abs(heat[c["heat_rows"]] - c["heat"]).max() <= parity.HEAT_ATOL
    dense = fe.debug_read("sp.dense").reshape(h // 8, w // 8, 256)
    assert np.abs(np.transpose(dense, (2, 0, 1))[:, ::8, ::8] - c["dense_desc_px"]).max() <= parity.DESC_ATOL


@pytest.mark.parametrize("n", [256, 512])
def test_lightglue_vs_cv2dnn_golden(fe, golden_dir, n):
    """The N0 x N1 log-assignment matrix after all nine layers, and the match list, against cv2.dnn's execution of
    lightglue_sim.onnx (TopK / mutual / filter applied in numpy to cv2.dnn's matrix by the fixture generator)."""
    c = np.load(os.path.join(golden_dir, f"cv2dnn_lg_synth_n{n}.npz"))
    k0, k1, d0, d1, _ = synth.lightglue_inputs(n, 200 + n)
    fe.debug_read("lg.S")                                    # arms the capture
    m, ms = fe.match(k0, k1, d0, d1, 480, 640)
    S = fe.debug_read("lg.S").reshape(n, n)
    assert np.abs(S[c["S_rows"]] - c["S"]).max() <= 5e-3 * max(1.0, np.abs(c["S"]).max() / 100)
    r = parity.compare_matches(c["matches"], c["mscores"], m, ms)
    assert r["only_ref"] == 0 and r["only_tst"] == 0

# ---- descriptor binarisation and L2 projection matching -------------------------------------------
def test_binarized_descriptors_bit_exact(fe):
    """8(f).1: the sampler's fused sign binarisation equals Frame::binarize_descriptors of the SAME descriptors (bit-exact),
    and the stand-alone entry point equals the oracle on arbitrary descriptors including +-0 and denormals."""
    from oracle import frontend_aux_ref as aux
    imgs = np.stack([synth.frame(s, 240, 320) for s in (5, 6)])
    fe.extract_device_from_host(imgs)
    for slot in range(2):
        k, s, d = fe.read_slot(slot)
        b = fe.read_slot_bin(slot)
        assert b.shape == (len(k), 256) and len(k) > 50
        assert np.array_equal(b, aux.binarize_descriptors(d))
    rng = np.random.RandomState(0)
    d = rng.randn(1000, 256).astype(np.float32)
    d[0, :4] = [0.0, -0.0, 1e-40, -1e-40]
    b, w = fe.binarize(d)
    assert np.array_equal(b, aux.binarize_descriptors(d))
    assert np.array_equal(w, aux.pack_bits(aux.binarize_descriptors(d)))
    b0, w0 = fe.binarize(np.zeros((0, 256), np.float32))
    assert b0.shape == (0, 256)


def test_l2_best2_vs_oracle(fe):
    """8(f).2: best / second-best L2 match over ragged candidate lists (empty lists, exact ties, every db row)."""
    from oracle import frontend_aux_ref as aux
    rng = np.random.RandomState(11)
    nq, nd = 300, 700
    q = rng.randn(nq, 256).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    db = rng.randn(nd, 256).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    db[17] = db[5]
    off, idx = [0], []
    for i in range(nq):
        k = 0 if i % 37 == 0 else (nd if i == 1 else rng.randint(1, 40))
        c = rng.choice(nd, size=k, replace=False).tolist()
        if i == 2:
            c = [17, 5, 100, 3]
        idx += c
        off.append(len(idx))
    off, idx = np.array(off, np.int32), np.array(idx, np.int32)
    b1, i1, b2, i2 = fe.l2_best2(q, db, off, idx)
    r1, j1, r2, j2 = aux.l2_best2(q, db, off, idx)
    assert np.allclose(b1, r1, rtol=2e-6, atol=1e-6) and np.allclose(b2, r2, rtol=2e-6, atol=1e-6)
    # indices identical except where two candidates are closer than the fp32 / fp64 accumulation difference
    gap_ok = np.abs(r2 - r1) > 1e-5
    assert np.array_equal(i1[gap_ok], j1[gap_ok])
    assert i1[0] == -1 and b1[0] == 256.0                   # empty candidate list
    assert i1[2] == 17 and i2[2] == 5                       # exact tie: list order wins (strict <)


def test_l2_best2_slots_equals_host_form(fe):
    """8(f).2, slot-resident: the same answer as rfe_l2_best2 on the descriptors read back from the slots, bit for bit."""
    imgs = np.stack(synth.frame_pair(90, 240, 320, shift=(5, 2)))
    fe.extract_device_from_host(imgs)
    (k0, _, d0), (k1, _, d1) = fe.read_slot(0), fe.read_slot(1)
    rng = np.random.RandomState(3)
    off, idx = [0], []
    for i in range(len(k0)):
        # candidates = keypoints of the other frame inside a 40-px window (the shape of SearchByProjection's GetFeaturesInArea)
        near = np.nonzero(np.abs(k1 - k0[i]).max(1) <= 20)[0]
        idx += rng.permutation(near).tolist()
        off.append(len(idx))
    off, idx = np.array(off, np.int32), np.array(idx, np.int32)
    a = fe.l2_best2_slots(0, 1, off, idx)
    b = fe.l2_best2(d0, d1, off, idx)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[1] >= 0).sum() > 50
    # host queries (MapPoint descriptors in SearchByProjection1) against the slot-resident frame
    q = d0[::3] + np.float32(0.01) * rng.randn(*d0[::3].shape).astype(np.float32)
    off3 = np.concatenate([[0], np.cumsum(np.diff(off)[::3])]).astype(np.int32)
    idx3 = np.concatenate([idx[off[i]:off[i + 1]] for i in range(0, len(k0), 3)] + [np.zeros(0, np.int32)]).astype(np.int32)
    a = fe.l2_best2_slots(None, 1, off3, idx3, q=q)
    b = fe.l2_best2(q, d1, off3, idx3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_one_to_many_with_layer0_cache_equals_batched_matching(fe):
    """8(f).3: one KeyFrame against several neighbours with the per-slot layer-0 cache == rfe_lg_match_slots_batch on the same
    pairs: the same match lists, scores within 1e-4 (the cache is built in a pass with fewer rows, whose GEMMs may use the
    other tiling: the lo-product accumulation order differs in the last bit, and nine layers amplify it to ~6e-6); the cache is reused on the second call,
    rebuilt after the slot is overwritten or when the normalisation size changes, and a stale entry is never used."""

    def same(a, b):
        # identical pairs except within the stated margin of the 0.1 filter, scores within the stated tolerance (tests/parity.py);
        # measured: identical lists, scores within 6e-6
        r = parity.compare_matches(a[0], a[1], b[0], b[1])
        return r["only_ref"] + r["only_tst"] <= 1 and r["mscore_maxabs"] <= 1e-4

    h, w = 240, 320
    imgs = np.stack([synth.frame_pair(100 + i, h, w, shift=(3 * i - 4, 2 * i - 3))[i % 2] for i in range(5)]
                    + [synth.frame_pair(100, h, w, shift=(6, -5))[1]])
    fe.extract_device_from_host(imgs)                       # slots 0..5
    others = [1, 2, 3, 5]
    fe.match_slots_batch([0] * 4, others, h, w)
    want = [fe.read_result(i) for i in range(4)]
    h0, b0 = fe.cache_stats()
    fe.match_one_to_many(0, others, h, w)
    got = [fe.read_result(i) for i in range(4)]
    h1, b1 = fe.cache_stats()
    assert b1 - b0 == 5 and h1 == h0                        # five distinct slots built, none found
    for w_, g_ in zip(want, got):
        assert same(w_, g_)
    assert max(len(m) for m, _ in want) > 30
    fe.match_one_to_many(0, [5, 3], h, w)                   # all three states come from the cache
    h2, b2 = fe.cache_stats()
    assert b2 == b1 and h2 - h1 == 3
    assert same(fe.read_result(0), want[3]) and same(fe.read_result(1), want[2])
    # other normalisation size (the KeyPoint overloads' 300 x 400): entries rebuilt, results follow
    fe.match_slots_batch([0, 0], [1, 2], 300, 400)
    w34 = [fe.read_result(i) for i in range(2)]
    fe.match_one_to_many(0, [1, 2], 300, 400)
    h3, b3 = fe.cache_stats()
    assert b3 - b2 == 3
    for i in range(2):
        assert same(fe.read_result(i), w34[i])
    # overwrite slot 2 with slot 5's features: its cache entry must be rebuilt, not reused
    k5, s5, d5 = fe.read_slot(5)
    fe.write_slot(2, k5, d5, s5)
    fe.match_one_to_many(0, [2], 300, 400)
    h4, b4 = fe.cache_stats()
    assert b4 - b3 == 1 and h4 - h3 == 1
    fe.match_slots_batch([0], [5], 300, 400)
    m_ref, s_ref = fe.read_result(0)
    fe.match_one_to_many(0, [2], 300, 400)
    assert same(fe.read_result(0), (m_ref, s_ref))
