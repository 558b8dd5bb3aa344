"""CPU tests of the section-8(f) helper oracles (binarisation, L2 best-two)."""
import os

import numpy as np

from oracle import frontend_aux_ref as aux


def test_binarize_known_answer():
    d = np.array([[0.5, -0.5, 0.0, -0.0, 1e-30, -1e-30] + [0.0] * 250], np.float32)
    b = aux.binarize_descriptors(d)
    assert b.dtype == np.uint8 and b.shape == (1, 256)
    assert b[0, :6].tolist() == [1, 0, 0, 0, 1, 0]          # strictly greater than zero (cv::THRESH_BINARY)
    assert aux.pack_bits(b)[0, 0] == 0b010001


def test_l2_best2_matches_brute_force():
    rng = np.random.RandomState(3)
    q = rng.randn(20, 256).astype(np.float32)
    db = rng.randn(50, 256).astype(np.float32)
    db[7] = db[3]                                            # an exact tie: the first candidate in list order must win
    off = [0]
    idx = []
    for i in range(20):
        k = rng.randint(0, 12)
        c = rng.choice(50, size=k, replace=False).tolist()
        if i == 4:
            c = [7, 3, 11]
        idx += c
        off.append(len(idx))
    b1, i1, b2, i2 = aux.l2_best2(q, db, np.array(off), np.array(idx, np.int64))
    for i in range(20):
        c = idx[off[i]:off[i + 1]]
        if not c:
            assert i1[i] == -1 and b1[i] == 256.0 and i2[i] == -1
            continue
        d = np.sqrt(((q[i][None] - db[c]).astype(np.float64) ** 2).sum(1)).astype(np.float32)
        order = np.argsort(d, kind="stable")
        assert i1[i] == c[order[0]] and np.isclose(b1[i], d[order[0]])
        if len(c) > 1:
            assert np.isclose(b2[i], d[order[1]])
    assert i1[4] == 7 and i2[4] == 3                         # tie broken by list order (strict <)


def test_adaptive_threshold_known_answers():
    """8(f).4: the reference's compiled-out adaptive score rule (superpoint_onnx.cc:192-210)."""
    from oracle import frontend_aux_ref as aux
    s = np.array([0.5, 0.25, 0.25, 0.0], np.float32)             # mean 0.25, variance 0.03125 (all exact in binary)
    sig = 0.02 / (1.0 + np.exp(-0.02 * (270.0 - 270.0)))         # lastmatch = 270: the logistic term is 0.01
    want = np.float32(0.25 - 0.6 * float(np.sqrt(np.float32(0.03125))) - sig)   # sqrt in float, the rest in double (C++ promotion)
    assert aux.adaptive_threshold(s, 270.0) == want
    assert list(aux.adaptive_filter(s, 270.0)) == [0, 1, 2]       # 0.0 < threshold (0.1339...) is dropped
    # the logistic term grows with the previous frame's match count: more matches -> lower threshold -> more keypoints
    t_lo, t_hi = aux.adaptive_threshold(s, 0.0), aux.adaptive_threshold(s, 1000.0)
    assert t_hi < t_lo and abs(float(t_lo - t_hi) - 0.02 * (1 / (1 + np.exp(-14.6)) - 1 / (1 + np.exp(5.4)))) < 1e-6
    rng = np.random.RandomState(0)
    sc = (rng.rand(2000).astype(np.float32) ** 4) * np.float32(0.6)
    keep = aux.adaptive_filter(sc, 150.0)
    thr = aux.adaptive_threshold(sc, 150.0)
    assert 0 < len(keep) <= 2000 and (sc[keep] >= thr).all() and (np.delete(sc, keep) < thr).all()


def test_algorithmic_flop_formulas_match_the_survey():
    """LightGlue 9*(4 980 736 N + 4096 N^2) + 263 168 N + 512 N^2 FLOP per pair = 14.0 / 32.9 / 85.4 / 249.1 GFLOP
    at N = 256 / 512 / 1024 / 2048, attention 9*4096 N^2 = 2.4 / 9.7 / 38.7 / 154.6; SuperPoint 52.10 GFLOP per 640x480 frame
    (sum of 2 k^2 Cin Cout Hout Wout over the twelve convolutions).  bench.py and tools/lg_sweep.py divide by these."""
    import bench
    lg = lambda n: 9 * (4980736 * n + 4096 * n * n) + 263168 * n + 512 * n * n
    assert [round(lg(n) / 1e9, 1) for n in (256, 512, 1024, 2048)] == [14.0, 32.9, 85.4, 249.1]
    assert [round(9 * 4096 * n * n / 1e9, 1) for n in (256, 512, 1024, 2048)] == [2.4, 9.7, 38.7, 154.6]
    h, w = 480, 640
    convs = [(3, 1, 64, 1), (3, 64, 64, 1), (3, 64, 64, 2), (3, 64, 64, 2), (3, 64, 128, 4), (3, 128, 128, 4), (3, 128, 128, 8),
             (3, 128, 128, 8), (3, 128, 256, 8), (1, 256, 65, 8), (3, 128, 256, 8), (1, 256, 256, 8)]     # (k, Cin, Cout, stride of the map)
    sp = sum(2 * k * k * ci * co * (h // s) * (w // s) for k, ci, co, s in convs)
    assert abs(sp - bench.SP_FLOPS_PER_FRAME) / sp < 2e-3, sp
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "lg_sweep.py")).read()
    assert "9 * (4980736 * n + 4096 * n * n) + 263168 * n + 512 * n * n" in src and "9 * 4096 * n * n" in src


def test_topk_keep_order():
    from oracle import frontend_aux_ref as aux
    s = np.float32([0.5, 0.9, 0.5, 0.1, 0.9, 0.5])
    assert aux.topk_keep_order(s, 0).tolist() == [0, 1, 2, 3, 4, 5]
    assert aux.topk_keep_order(s, 9).tolist() == [0, 1, 2, 3, 4, 5]
    assert aux.topk_keep_order(s, 2).tolist() == [1, 4]
    assert aux.topk_keep_order(s, 3).tolist() == [0, 1, 4]           # ties at the cut: the earliest 0.5 stays
    assert aux.topk_keep_order(s, 4).tolist() == [0, 1, 2, 4]
