"""CPU restatement of the host-side descriptor helpers next to the learned front end.
TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product.

  binarize_descriptors  <- Frame::binarize_descriptors     src/Frame.cc:1034-1043, KeyFrame.cc:113-123
  l2_best2              <- the candidate loop of SPmatcher::SearchByProjection*   src/Matchers/SPmatcher.cc:1225-1250
                           with DescriptorDistance_sp = cv::norm(a, b, NORM_L2)   src/Matchers/SPmatcher.cc:2184-2189
Parity: the reference has no test or fixture for these functions ("parity unpinned"); both are a few lines of host C++ and
are restated literally.
"""
import numpy as np


def binarize_descriptors(desc: np.ndarray) -> np.ndarray:
    """cv::threshold(row, tmp, 0, 1, THRESH_BINARY) then uchar(tmp): 1 where the element is > 0, else 0."""
    return (np.asarray(desc, np.float32) > 0).astype(np.uint8)


def pack_bits(b: np.ndarray) -> np.ndarray:
    """[N,256] 0/1 -> [N,8] uint32, bit j%32 of word j//32 = element j."""
    w = b.reshape(len(b), 8, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)
    return w.sum(axis=2).astype(np.uint32)


def descriptor_distance_sp(a: np.ndarray, b: np.ndarray) -> np.float32:
    """cv::norm(a, b, NORM_L2) on CV_32F rows: float difference, double accumulation, sqrt, cast to float."""
    d = (np.asarray(a, np.float32) - np.asarray(b, np.float32)).astype(np.float64)
    return np.float32(np.sqrt(np.sum(d * d)))


def l2_best2(q, db, cand_off, cand_idx, init_dist=256.0):
    """For every query: walk its candidate list in order; `if dist < best: second = best; best = dist elif dist < second:
    second = dist` with both starting at init_dist (SPmatcher.cc:1211-1250).  Returns best/second distance and index."""
    nq = len(q)
    b1 = np.full(nq, init_dist, np.float32)
    b2 = np.full(nq, init_dist, np.float32)
    i1 = np.full(nq, -1, np.int32)
    i2 = np.full(nq, -1, np.int32)
    for i in range(nq):
        for c in range(cand_off[i], cand_off[i + 1]):
            idx = int(cand_idx[c])
            dist = descriptor_distance_sp(q[i], db[idx])
            if dist < b1[i]:
                b2[i], i2[i] = b1[i], i1[i]
                b1[i], i1[i] = dist, idx
            elif dist < b2[i]:
                b2[i], i2[i] = dist, idx
    return b1, i1, b2, i2


def adaptive_threshold(scores: np.ndarray, lastmatch: float) -> np.float32:
    """the score filter the reference compiles out (src/Extractors/superpoint_onnx.cc:192-210,
    `bool adaptivethresold = false`): threshold = mean - 0.6*sqrt(var) - 0.02 / (1 + exp(-0.02 (lastmatch - 270))).
    float32 sum / mean / variance accumulated in index order, the last expression in double, narrowed to float32."""
    s = np.asarray(scores, np.float32)
    n = len(s)
    acc = np.float32(0)
    for v in s:                                   # :196-199
        acc = np.float32(acc + v)
    mean = np.float32(acc / np.float32(n))
    var = np.float32(0)
    for v in s:                                   # :202-206
        d = np.float32(v - mean)
        var = np.float32(var + np.float32(d * d))
    var = np.float32(var / np.float32(n))
    return np.float32(float(mean) - 0.6 * float(np.sqrt(var)) - 0.02 / (1.0 + np.exp(-0.02 * (float(lastmatch) - 270.0))))


def adaptive_filter(scores: np.ndarray, lastmatch: float) -> np.ndarray:
    """Indices Extractor_PostProcess keeps under the adaptive rule: `if (scores[i] < threshold) continue;` (:226)."""
    s = np.asarray(scores, np.float32)
    return np.nonzero(~(s < adaptive_threshold(s, lastmatch)))[0]


def topk_keep_order(scores: np.ndarray, k: int) -> np.ndarray:
    """SURVEY.md 8(f).4 -- the `nfeatures` cap the reference stores and never applies (SPextractor.cc:84-146): indices of the k
    highest scores, returned in ascending (= the graph's row-major keypoint) order; among equal scores the earlier one stays
    (stable sort).  k <= 0 or k >= len(scores): everything."""
    s = np.asarray(scores, np.float32)
    if k <= 0 or k >= len(s):
        return np.arange(len(s))
    order = np.argsort(-s.astype(np.float64), kind="stable")[:k]
    return np.sort(order)
