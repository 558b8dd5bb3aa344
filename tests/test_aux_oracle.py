"""CPU tests of the section-8(f) helper oracles (binarisation, L2 best-two)."""
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
