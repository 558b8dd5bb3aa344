"""ctypes binding of include/rover_fe.h (the same stub a reference-side Python harness would use)."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
DESC_DIM = 256

RFE_OK, RFE_ERR_INVALID, RFE_ERR_CUDA, RFE_ERR_IO, RFE_ERR_CAPACITY, RFE_ERR_NO_DEVICE = range(6)


class RoverFeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rover_fe error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.environ.get("ROVER_FE_LIB", os.path.join(_HERE, "librover_fe.so"))


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("weights_path", C.c_char_p),
                ("max_batch", C.c_int), ("max_height", C.c_int), ("max_width", C.c_int),
                ("max_keypoints", C.c_int), ("flags", C.c_int)]

RFE_FLAG_NO_MATCHER, RFE_FLAG_NO_EXTRACTOR = 1, 2


_lib = None


def exported_symbols() -> list:
    """Every function include/rover_fe.h declares (parsed from the header)."""
    with open(os.path.join(ROOT, "include", "rover_fe.h")) as f:
        src = f.read()
    src = re.sub(r"#ifdef RFE_ENABLE_PROBES.*?#endif", "", src, flags=re.S)      # probe builds only (make PROBES=1)
    return sorted(set(re.findall(r"\b(rfe_[a-z0-9_]+)\s*\(", src)))


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise RoverFeError(-1, f"{p} not found: build it with `make` (python -c 'import __graft_entry__ as g; g.build()'); "
                               "there is no CPU fallback")
    lib = C.CDLL(p)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    P = C.POINTER
    lib.rfe_create.argtypes = [P(_Config), P(vp)]
    lib.rfe_destroy.argtypes = [vp]
    lib.rfe_destroy.restype = None
    lib.rfe_last_error.restype = C.c_char_p
    lib.rfe_sync.argtypes = [vp]
    lib.rfe_sp_extract_u8.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, ci]
    lib.rfe_sp_extract_device.argtypes = [vp, vp, ci, ci, ci, ci]
    lib.rfe_sp_set_topk.argtypes = [vp, ci]
    lib.rfe_set_sm_limit.argtypes = [vp, ci]
    lib.rfe_set_fast_mode.argtypes = [vp, ci]
    lib.rfe_sp_read_slot.argtypes = [vp, ci, vp, vp, vp, vp, ci]
    lib.rfe_sp_read_slot_bin.argtypes = [vp, ci, vp, vp, ci]
    lib.rfe_sp_write_slot.argtypes = [vp, ci, vp, vp, vp, ci]
    lib.rfe_binarize_descriptors.argtypes = [vp, vp, ci, vp, vp]
    lib.rfe_l2_best2.argtypes = [vp, vp, ci, vp, ci, vp, vp, cf, vp, vp, vp, vp]
    lib.rfe_l2_best2_slots.argtypes = [vp, vp, ci, ci, ci, vp, vp, cf, vp, vp, vp, vp]
    lib.rfe_lg_match.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci, ci, cf, vp, vp, vp]
    lib.rfe_lg_match_normalized.argtypes = [vp, vp, ci, vp, ci, vp, vp, cf, vp, vp, vp]
    lib.rfe_lg_match_slots.argtypes = [vp, ci, ci, ci, ci, cf, ci]
    lib.rfe_lg_match_slots_batch.argtypes = [vp, ci, vp, vp, ci, ci, cf]
    lib.rfe_lg_match_one_to_many.argtypes = [vp, ci, vp, ci, ci, ci, cf]
    lib.rfe_lg_cache_stats.argtypes = [vp, P(C.c_longlong), P(C.c_longlong)]
    lib.rfe_match_pairs_u8.argtypes = [vp, vp, ci, ci, ci, ci, cf, vp, vp, vp, vp, vp, ci]
    lib.rfe_pairs_submit.argtypes = [vp, vp, ci, ci, ci, ci]
    lib.rfe_pairs_collect.argtypes = [vp, cf, vp, vp, vp, vp, vp, ci]
    lib.rfe_pairs_collect_begin.argtypes = [vp, cf, vp, vp, vp, ci]
    lib.rfe_pairs_collect_end.argtypes = [vp, vp, vp]
    lib.rfe_pairs_collect_begin_full.argtypes = [vp, cf, vp, vp, vp, vp, vp, ci]
    lib.rfe_alloc_pinned.argtypes = [C.c_size_t]
    lib.rfe_alloc_pinned.restype = vp
    lib.rfe_free_pinned.argtypes = [vp]
    lib.rfe_free_pinned.restype = None
    lib.rfe_lg_copy_results_device.argtypes = [vp, ci, vp, vp, vp]
    lib.rfe_lg_read_result.argtypes = [vp, ci, vp, vp, vp, ci]
    lib.rfe_get_timer_ms.argtypes = [vp, C.c_char_p]
    lib.rfe_get_timer_ms.restype = C.c_double
    lib.rfe_transfer_bytes.argtypes = [vp, P(C.c_ulonglong), P(C.c_ulonglong)]
    lib.rfe_kernel_launches.argtypes = [vp]
    lib.rfe_kernel_launches.restype = C.c_longlong
    lib.rfe_profile.argtypes = [vp, ci]
    lib.rfe_profile_select.argtypes = [vp, C.c_char_p]
    lib.rfe_profile_read.argtypes = [vp, C.c_char_p, P(C.c_double), P(C.c_longlong), ci]
    lib.rfe_debug_read.argtypes = [vp, C.c_char_p, vp, C.c_size_t, P(C.c_size_t)]
    lib.rfe_debug_gemm.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci]
    if hasattr(lib, "rfe_debug_probe"):
        lib.rfe_debug_probe.argtypes = [vp, ci, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class FrontEnd:
    """One rfe_ctx.  Mirrors what the reference's SuperPointOnnxRunner + LightGlueDecoupleOnnxRunner pair does."""

    def __init__(self, device: int = 0, stream: int | None = None, weights: str | None = None, max_batch: int = 8,
                 max_height: int = 480, max_width: int = 768, max_keypoints: int = 4096, flags: int = 0):
        self.lib = load_library()
        weights = weights or os.environ.get("ROVER_FE_WEIGHTS") or os.path.join(ROOT, "weights", "rover_fe.rfw")
        cfg = _Config(device, stream, weights.encode(), max_batch, max_height, max_width, max_keypoints, flags)
        h = C.c_void_p()
        self.ctx = None
        self._check(self.lib.rfe_create(C.byref(cfg), C.byref(h)))
        self.ctx = h
        self.cap = max_keypoints
        self.max_batch = max_batch

    def _check(self, code, allow=()):
        if code != RFE_OK and code not in allow:
            raise RoverFeError(code, (self.lib.rfe_last_error() or b"").decode())
        return code

    def close(self):
        if self.ctx is not None:
            self.lib.rfe_destroy(self.ctx)
            self.ctx = None
            self._mp_sets = {}
            for p in getattr(self, "_pinned", []):
                self.lib.rfe_free_pinned(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.rfe_sync(self.ctx))

    # ---- SuperPoint -----------------------------------------------------------------------------
    def extract(self, images: np.ndarray, want_desc: bool = True):
        """images: uint8 [H,W] or [B,H,W] (host).  Returns a list of (kpts int32 [N,2] xy, scores [N], desc [N,256])."""
        imgs = np.ascontiguousarray(images if images.ndim == 3 else images[None], dtype=np.uint8)
        b, h, w = imgs.shape
        cap = self.cap
        kp = np.empty((b, cap, 2), np.int32)
        sc = np.empty((b, cap), np.float32)
        de = np.empty((b, cap, DESC_DIM), np.float32) if want_desc else None
        cnt = np.zeros(b, np.int32)
        self._check(self.lib.rfe_sp_extract_u8(self.ctx, _ptr(imgs), h, w, w, b, _ptr(kp), _ptr(sc), _ptr(de), _ptr(cnt), cap))
        return [(kp[i, :cnt[i]].copy(), sc[i, :cnt[i]].copy(), de[i, :cnt[i]].copy() if want_desc else None)
                for i in range(b)]

    def set_topk(self, k: int):
        """Keep the k best keypoints per image in later extractions (k <= 0: all of them, the reference's behaviour)."""
        self._check(self.lib.rfe_sp_set_topk(self.ctx, int(k)))

    def set_fast_mode(self, on: bool):
        """Labelled fast mode (NOT the parity path): hi-only fp16 MMAs in SuperPoint's 3x3 convolutions."""
        self._check(self.lib.rfe_set_fast_mode(self.ctx, 1 if on else 0))

    def set_sm_limit(self, max_sms: int):
        """Persistent kernels use at most max_sms SMs (0 = all): leaves SMs to communication kernels running beside them."""
        self._check(self.lib.rfe_set_sm_limit(self.ctx, int(max_sms)))

    def extract_device(self, d_ptr: int, h: int, w: int, stride: int, batch: int):
        self._check(self.lib.rfe_sp_extract_device(self.ctx, C.c_void_p(d_ptr), h, w, stride, batch))

    def extract_device_from_host(self, images: np.ndarray):
        """Upload with torch (plumbing only) and run the device-resident extractor; features stay in the slots."""
        import torch
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        b, h, w = imgs.shape
        self._dev_imgs = torch.from_numpy(imgs).cuda()
        torch.cuda.synchronize()
        self.extract_device(self._dev_imgs.data_ptr(), h, w, w, b)
        self.sync()

    def read_slot(self, slot: int, want_desc: bool = True):
        cap = self.cap
        kp = np.empty((cap, 2), np.int32)
        sc = np.empty(cap, np.float32)
        de = np.empty((cap, DESC_DIM), np.float32) if want_desc else None
        n = C.c_int(0)
        self._check(self.lib.rfe_sp_read_slot(self.ctx, slot, _ptr(kp), _ptr(sc), _ptr(de), C.byref(n), cap))
        n = n.value
        return kp[:n].copy(), sc[:n].copy(), de[:n].copy() if want_desc else None

    def write_slot(self, slot: int, kpts_xy, desc, scores=None):
        """Upload host features (e.g. a stored KeyFrame's) into a device slot for match_slots / match_slots_batch."""
        kp = np.ascontiguousarray(np.rint(np.asarray(kpts_xy)), np.int32).reshape(-1, 2)
        de = np.ascontiguousarray(desc, np.float32).reshape(-1, DESC_DIM)
        if len(kp) != len(de):
            raise ValueError("keypoints and descriptors differ in length")
        sc = None if scores is None else np.ascontiguousarray(scores, np.float32).reshape(-1)
        if sc is not None and len(sc) != len(kp):
            raise ValueError("scores and keypoints differ in length")
        self._check(self.lib.rfe_sp_write_slot(self.ctx, slot, _ptr(kp), _ptr(sc), _ptr(de), len(kp)))

    def read_slot_bin(self, slot: int) -> np.ndarray:
        """Sign-binarised descriptors [N,256] uint8 (0/1) of a feature slot (Frame::binarize_descriptors)."""
        out = np.empty((self.cap, DESC_DIM), np.uint8)
        n = C.c_int(0)
        self._check(self.lib.rfe_sp_read_slot_bin(self.ctx, slot, _ptr(out), C.byref(n), self.cap))
        return out[:n.value].copy()

    def binarize(self, desc: np.ndarray):
        """(bin uint8 [N,256] 0/1, bits uint32 [N,8]) of host descriptors [N,256]."""
        d = np.ascontiguousarray(desc, np.float32).reshape(-1, DESC_DIM)
        n = len(d)
        b = np.empty((n, DESC_DIM), np.uint8)
        w = np.empty((n, 8), np.uint32)
        self._check(self.lib.rfe_binarize_descriptors(self.ctx, _ptr(d), n, _ptr(b), _ptr(w)))
        return b, w

    def l2_best2(self, q: np.ndarray, db: np.ndarray, cand_off: np.ndarray, cand_idx: np.ndarray, init_dist: float = 256.0):
        """Best / second-best L2 match of every query over its candidate list (SearchByProjection inner loop).
        Returns (best_dist, best_idx, second_dist, second_idx)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, DESC_DIM)
        db = np.ascontiguousarray(db, np.float32).reshape(-1, DESC_DIM)
        off = np.ascontiguousarray(cand_off, np.int32)
        idx = np.ascontiguousarray(cand_idx, np.int32)
        nq = len(q)
        b1, b2 = np.empty(nq, np.float32), np.empty(nq, np.float32)
        i1, i2 = np.empty(nq, np.int32), np.empty(nq, np.int32)
        self._check(self.lib.rfe_l2_best2(self.ctx, _ptr(q), nq, _ptr(db), len(db), _ptr(off), _ptr(idx), init_dist,
                                          _ptr(b1), _ptr(i1), _ptr(b2), _ptr(i2)))
        return b1, i1, b2, i2

    def l2_best2_slots(self, q_slot, db_slot: int, cand_off: np.ndarray, cand_idx: np.ndarray, init_dist: float = 256.0, q=None):
        """l2_best2 with the database (and, when q is None, the queries) taken from device-resident feature slots."""
        qh = None if q is None else np.ascontiguousarray(q, np.float32).reshape(-1, DESC_DIM)
        off = np.ascontiguousarray(cand_off, np.int32)
        idx = np.ascontiguousarray(cand_idx, np.int32)
        nq = len(off) - 1
        b1, b2 = np.empty(nq, np.float32), np.empty(nq, np.float32)
        i1, i2 = np.empty(nq, np.int32), np.empty(nq, np.int32)
        self._check(self.lib.rfe_l2_best2_slots(self.ctx, _ptr(qh), -1 if q_slot is None else q_slot, db_slot, nq, _ptr(off), _ptr(idx), init_dist,
                                                _ptr(b1), _ptr(i1), _ptr(b2), _ptr(i2)))
        return b1, i1, b2, i2

    # ---- LightGlue -------------------------------------------------------------------------------
    def match(self, kpts0, kpts1, desc0, desc1, norm_h: int, norm_w: int, thresh: float = 0.0):
        """Pixel keypoints [N,2] (x,y), descriptors [N,256].  Returns (matches int32 [K,2], mscores [K])."""
        k0 = np.ascontiguousarray(kpts0, np.float32).reshape(-1, 2)
        k1 = np.ascontiguousarray(kpts1, np.float32).reshape(-1, 2)
        d0 = np.ascontiguousarray(desc0, np.float32).reshape(-1, DESC_DIM)
        d1 = np.ascontiguousarray(desc1, np.float32).reshape(-1, DESC_DIM)
        n0, n1 = len(k0), len(k1)
        m = np.empty((max(n0, 1), 2), np.int32)
        s = np.empty(max(n0, 1), np.float32)
        k = C.c_int(0)
        self._check(self.lib.rfe_lg_match(self.ctx, _ptr(k0), n0, _ptr(k1), n1, _ptr(d0), _ptr(d1), norm_h, norm_w,
                                          thresh, _ptr(m), _ptr(s), C.byref(k)))
        return m[:k.value].copy(), s[:k.value].copy()

    def match_normalized(self, kn0, kn1, desc0, desc1, thresh: float = 0.0):
        """Keypoints already normalised by the caller (NormalizeKeypoints, transform.cpp:19-32): what the reference's
        Matcher_Inference receives."""
        k0 = np.ascontiguousarray(kn0, np.float32).reshape(-1, 2)
        k1 = np.ascontiguousarray(kn1, np.float32).reshape(-1, 2)
        d0 = np.ascontiguousarray(desc0, np.float32).reshape(-1, DESC_DIM)
        d1 = np.ascontiguousarray(desc1, np.float32).reshape(-1, DESC_DIM)
        n0, n1 = len(k0), len(k1)
        m = np.empty((max(n0, 1), 2), np.int32)
        s = np.empty(max(n0, 1), np.float32)
        k = C.c_int(0)
        self._check(self.lib.rfe_lg_match_normalized(self.ctx, _ptr(k0), n0, _ptr(k1), n1, _ptr(d0), _ptr(d1), thresh,
                                                     _ptr(m), _ptr(s), C.byref(k)))
        return m[:k.value].copy(), s[:k.value].copy()

    def match_slots(self, slot0: int, slot1: int, norm_h: int, norm_w: int, thresh: float = 0.0, rslot: int = 0):
        self._check(self.lib.rfe_lg_match_slots(self.ctx, slot0, slot1, norm_h, norm_w, thresh, rslot))

    def match_slots_batch(self, slots0, slots1, norm_h: int, norm_w: int, thresh: float = 0.0):
        s0 = np.ascontiguousarray(slots0, np.int32)
        s1 = np.ascontiguousarray(slots1, np.int32)
        self._check(self.lib.rfe_lg_match_slots_batch(self.ctx, len(s0), _ptr(s0), _ptr(s1), norm_h, norm_w, thresh))

    def match_one_to_many(self, slot: int, others, norm_h: int, norm_w: int, thresh: float = 0.0):
        """Slot `slot` against every slot of `others` (result slot i = pair i), layer-0 state of every slot cached."""
        o = np.ascontiguousarray(others, np.int32)
        self._check(self.lib.rfe_lg_match_one_to_many(self.ctx, slot, _ptr(o), len(o), norm_h, norm_w, thresh))

    def cache_stats(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._check(self.lib.rfe_lg_cache_stats(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def match_pairs(self, images: np.ndarray, thresh: float = 0.0, want_kpts: bool = True):
        """images: uint8 [2*P, H, W] host array, pair p = images 2p, 2p+1.  One call = extract all + match all.
        Returns (list of keypoint arrays [N,2] per image, list of (matches [K,2], mscores [K]) per pair)."""
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        b, h, w = imgs.shape
        npairs, cap = b // 2, self.cap
        if not hasattr(self, "_mp_buf") or self._mp_buf[0].shape[0] != b:
            self._mp_buf = (np.empty((b, cap, 2), np.int32), np.zeros(b, np.int32), np.empty((npairs, cap, 2), np.int32),
                            np.empty((npairs, cap), np.float32), np.zeros(npairs, np.int32))
        kp, kc, m, ms, mc = self._mp_buf
        self._check(self.lib.rfe_match_pairs_u8(self.ctx, _ptr(imgs), h, w, w, npairs, thresh, _ptr(kp) if want_kpts else None,
                                                _ptr(kc), _ptr(m), _ptr(ms), _ptr(mc), cap))
        kpts = [kp[i, :kc[i]] for i in range(b)] if want_kpts else None
        return kpts, [(m[i, :mc[i]], ms[i, :mc[i]]) for i in range(npairs)]

    def pairs_submit(self, images: np.ndarray):
        """Pipelined matching, step 1: images uint8 [2*P, H, W] (pinned host memory for an asynchronous copy).  The array
        must stay alive until the matching pairs_collect()."""
        b, h, w = images.shape
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append(images)
        self._check(self.lib.rfe_pairs_submit(self.ctx, _ptr(images), h, w, w, b // 2))

    def pairs_collect(self, thresh: float = 0.0, want_kpts: bool = True):
        """Pipelined matching, step 2: results of the oldest submitted batch, same form as match_pairs()."""
        imgs = self._inflight.pop(0)
        b = imgs.shape[0]
        npairs, cap = b // 2, self.cap
        if not hasattr(self, "_mp_buf") or self._mp_buf[0].shape[0] != b:
            self._mp_buf = (np.empty((b, cap, 2), np.int32), np.zeros(b, np.int32), np.empty((npairs, cap, 2), np.int32),
                            np.empty((npairs, cap), np.float32), np.zeros(npairs, np.int32))
        kp, kc, m, ms, mc = self._mp_buf
        self._check(self.lib.rfe_pairs_collect(self.ctx, thresh, _ptr(kp) if want_kpts else None, _ptr(kc), _ptr(m), _ptr(ms),
                                               _ptr(mc), cap))
        kpts = [kp[i, :kc[i]] for i in range(b)] if want_kpts else None
        return kpts, [(m[i, :mc[i]], ms[i, :mc[i]]) for i in range(npairs)]

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked host memory (rfe_alloc_pinned); freed with the FrontEnd."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.lib.rfe_alloc_pinned(n)
        if not p:
            raise RoverFeError(RFE_ERR_CUDA, (self.lib.rfe_last_error() or b"").decode())
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (C.c_uint8 * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def pairs_collect_begin(self, thresh: float = 0.0, want_kpts: bool = True, want_desc: bool = False):
        """First half of pairs_collect(): enqueue the matcher and the result copies of the oldest submitted batch.
        want_desc: also return scores and fp32 descriptors of every image (what SPextractor::operator() returns)."""
        imgs = self._inflight.pop(0)
        b = imgs.shape[0]
        npairs, cap = b // 2, self.cap
        # two result buffer sets (pinned): the arrays handed out by the previous collect stay valid while this one is filled
        self._mp_sets = getattr(self, "_mp_sets", {})
        self._mp_flip = 1 - getattr(self, "_mp_flip", 0)
        key = (b, self._mp_flip, want_desc)
        if key not in self._mp_sets:
            pe = self.pinned_empty
            self._mp_sets[key] = (pe((b, cap, 2), np.int32), np.zeros(b, np.int32), pe((npairs, cap, 2), np.int32),
                                  pe((npairs, cap), np.float32), np.zeros(npairs, np.int32),
                                  pe((b, cap), np.float32) if want_desc else None, pe((b, cap, DESC_DIM), np.float32) if want_desc else None)
        self._collecting = (self._mp_sets[key], b, npairs, want_kpts)
        kp, kc, m, ms, mc, sc, de = self._mp_sets[key]
        self._check(self.lib.rfe_pairs_collect_begin_full(self.ctx, thresh, _ptr(kp) if want_kpts else None, _ptr(sc), _ptr(de),
                                                          _ptr(m), _ptr(ms), cap))

    def pairs_collect_end(self):
        """Returns (keypoints per image, (matches, mscores) per pair); after pairs_collect_begin(want_desc=True) a third element:
        (scores, descriptors) per image."""
        (kp, kc, m, ms, mc, sc, de), b, npairs, want_kpts = self._collecting
        self._check(self.lib.rfe_pairs_collect_end(self.ctx, _ptr(kc), _ptr(mc)))
        kpts = [kp[i, :kc[i]] for i in range(b)] if want_kpts else None
        res = [(m[i, :mc[i]], ms[i, :mc[i]]) for i in range(npairs)]
        if de is None:
            return kpts, res
        return kpts, res, [(sc[i, :kc[i]], de[i, :kc[i]]) for i in range(b)]

    def copy_results_device(self, n_pairs: int, d_matches: int, d_mscores: int, d_counts: int):
        """Device pointers (e.g. torch tensors' data_ptr()) receive the match results of the last batched match."""
        self._check(self.lib.rfe_lg_copy_results_device(self.ctx, n_pairs, C.c_void_p(d_matches), C.c_void_p(d_mscores),
                                                        C.c_void_p(d_counts)))

    def read_result(self, rslot: int = 0):
        m = np.empty((self.cap, 2), np.int32)
        s = np.empty(self.cap, np.float32)
        k = C.c_int(0)
        self._check(self.lib.rfe_lg_read_result(self.ctx, rslot, _ptr(m), _ptr(s), C.byref(k), self.cap))
        return m[:k.value].copy(), s[:k.value].copy()

    # ---- introspection -----------------------------------------------------------------------------
    def timer_ms(self, name: str) -> float:
        return float(self.lib.rfe_get_timer_ms(self.ctx, name.encode()))

    def transfer_bytes(self):
        """(host->device, device->host) bytes copied by the pair-matching path so far."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self._check(self.lib.rfe_transfer_bytes(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_launches(self) -> int:
        return int(self.lib.rfe_kernel_launches(self.ctx))

    def profile(self, enable: bool = True, select: str | None = None):
        """Per-launch CUDA-event timing; `select` restricts it to tags with that prefix."""
        self._check(self.lib.rfe_profile_select(self.ctx, None if select is None else select.encode()))
        self._check(self.lib.rfe_profile(self.ctx, 1 if enable else 0))

    def profile_read(self, prefix: str | None = None, reset: bool = False):
        """(total_ms, launches) of the profiled kernels whose tag starts with `prefix`."""
        ms, n = C.c_double(0), C.c_longlong(0)
        self._check(self.lib.rfe_profile_read(self.ctx, None if prefix is None else prefix.encode(), C.byref(ms),
                                              C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    def debug_read(self, name: str, shape=None) -> np.ndarray:
        nb = C.c_size_t(0)
        self._check(self.lib.rfe_debug_read(self.ctx, name.encode(), None, 0, C.byref(nb)))
        out = np.empty(nb.value // 4, np.float32)
        if nb.value:
            self._check(self.lib.rfe_debug_read(self.ctx, name.encode(), _ptr(out), nb.value, C.byref(nb)))
        return out.reshape(shape) if shape is not None else out

    def debug_probe_shift(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float32).reshape(136, 64)
        b = np.ascontiguousarray(b, np.float32).reshape(64, 64)
        out = np.empty((9, 3, 128, 64), np.float32)
        self._check(self.lib.rfe_debug_probe(self.ctx, 0, _ptr(a), _ptr(b), _ptr(out)))
        return out

    def debug_gemm(self, a: np.ndarray, b: np.ndarray, bias: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        m, k = a.shape
        n = b.shape[0]
        d = np.empty((m, n), np.float32)
        bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
        self._check(self.lib.rfe_debug_gemm(self.ctx, _ptr(a), _ptr(b), _ptr(bias), _ptr(d), m, n, k))
        return d
