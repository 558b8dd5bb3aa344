"""Shared parity comparators.  The STATED TOLERANCES of this project live here.

fp32-equivalent mode (the product's default; SURVEY.md 8(d), calibrated in DESIGN.md):
  heat-map             abs diff <= HEAT_ATOL
  keypoint set         identical, except keypoints whose score is within SCORE_MARGIN of the 0.0005
                       threshold or that lose/win a 9x9 NMS comparison by < SCORE_MARGIN
  scores               abs diff <= SCORE_ATOL on common keypoints
  descriptors          max-abs diff <= DESC_ATOL on common keypoints
  matches              identical pairs, except pairs whose mscore is within MSCORE_MARGIN of the 0.1 filter
  mscores              abs diff <= MSCORE_ATOL on common pairs
Integer outputs (coordinates, match indices, ordering) are bit-exact.
"""
from __future__ import annotations

import numpy as np

HEAT_ATOL = 1e-5
SCORE_MARGIN = 1e-5
SCORE_ATOL = 1e-5
DESC_ATOL = 2e-5
MSCORE_MARGIN = 5e-3
MSCORE_ATOL = 5e-3
THRESHOLD = 0.0005
FILTER = 0.1


def compare_keypoints(k_ref, s_ref, k_tst, s_tst, strict=False, heat=None):
    """k_*: [N,2] int (x,y) row-major ordered; s_*: [N].  Returns dict with common index arrays.

    strict=True: the two sets must be identical (the goldens: DESIGN.md section 4 reports identical sets).
    Otherwise a keypoint present in only one set is accepted only if
      * its score is within SCORE_MARGIN of the 0.0005 threshold, or
      * `heat` (a heat-map [H,W] of either side) is given and the 9x9 neighbourhood of the point really holds a rival
        whose score is within SCORE_MARGIN of the point's (an NMS near-tie: `==` against the 9x9 maximum on exact floats).
    Anything else raises AssertionError: there is no allowance for unexplained differences.
    """
    k_ref, k_tst = np.asarray(k_ref).astype(np.int64), np.asarray(k_tst).astype(np.int64)
    key_r = k_ref[:, 1] * 100000 + k_ref[:, 0]
    key_t = k_tst[:, 1] * 100000 + k_tst[:, 0]
    assert np.all(np.diff(key_r) > 0), "reference keypoints not in row-major order"
    assert np.all(np.diff(key_t) > 0), "keypoints not in row-major (y, then x) order"
    common, ir, it = np.intersect1d(key_r, key_t, return_indices=True)
    only_r = np.setdiff1d(np.arange(len(key_r)), ir)
    only_t = np.setdiff1d(np.arange(len(key_t)), it)
    if strict:
        assert len(only_r) == 0 and len(only_t) == 0, (len(only_r), len(only_t))

    def explained(k, s):
        if abs(float(s) - THRESHOLD) < SCORE_MARGIN:
            return True
        if heat is None:
            return False
        x, y = int(k[0]), int(k[1])
        win = np.asarray(heat)[max(0, y - 4):y + 5, max(0, x - 4):x + 5].astype(np.float64).copy()
        win[min(4, y), min(4, x)] = -np.inf               # the point itself
        return bool((np.abs(win - float(np.asarray(heat)[y, x])) < SCORE_MARGIN).any())

    bad = [("ref", tuple(k_ref[i]), float(np.asarray(s_ref)[i])) for i in only_r if not explained(k_ref[i], np.asarray(s_ref)[i])]
    bad += [("tst", tuple(k_tst[i]), float(np.asarray(s_tst)[i])) for i in only_t if not explained(k_tst[i], np.asarray(s_tst)[i])]
    assert not bad, f"keypoints differ without a threshold margin or an NMS near-tie to explain it: {bad[:8]}"
    ds = np.abs(np.asarray(s_ref)[ir] - np.asarray(s_tst)[it])
    assert ds.size == 0 or ds.max() <= SCORE_ATOL, f"score diff {ds.max()}"
    return {"ref_idx": ir, "tst_idx": it, "only_ref": only_r, "only_tst": only_t,
            "score_maxabs": float(ds.max()) if ds.size else 0.0}


def compare_descriptors(d_ref, d_tst, atol=DESC_ATOL):
    d = np.abs(np.asarray(d_ref, dtype=np.float64) - np.asarray(d_tst, dtype=np.float64))
    m = float(d.max()) if d.size else 0.0
    assert m <= atol, f"descriptor max-abs diff {m} > {atol}"
    return m


def compare_matches(m_ref, s_ref, m_tst, s_tst):
    m_ref, m_tst = np.asarray(m_ref).astype(np.int64).reshape(-1, 2), np.asarray(m_tst).astype(np.int64).reshape(-1, 2)
    assert np.all(np.diff(m_tst[:, 0]) > 0), "matches not ascending in query index"
    key_r = m_ref[:, 0] * 100000 + m_ref[:, 1]
    key_t = m_tst[:, 0] * 100000 + m_tst[:, 1]
    common, ir, it = np.intersect1d(key_r, key_t, return_indices=True)
    only_r = np.setdiff1d(np.arange(len(key_r)), ir)
    only_t = np.setdiff1d(np.arange(len(key_t)), it)
    bad_r = np.abs(np.asarray(s_ref)[only_r] - FILTER) >= MSCORE_MARGIN
    bad_t = np.abs(np.asarray(s_tst)[only_t] - FILTER) >= MSCORE_MARGIN
    assert bad_r.sum() + bad_t.sum() == 0, (
        f"match sets differ beyond the filter margin: only_ref={len(only_r)} only_test={len(only_t)}")
    ds = np.abs(np.asarray(s_ref)[ir] - np.asarray(s_tst)[it])
    assert ds.size == 0 or ds.max() <= MSCORE_ATOL, f"mscore diff {ds.max()}"
    return {"common": len(common), "only_ref": len(only_r), "only_tst": len(only_t),
            "mscore_maxabs": float(ds.max()) if ds.size else 0.0}
