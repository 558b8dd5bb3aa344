"""Numerics of the operand formats the tensor-core kernels use, restated in numpy (np.float16 conversion = cvt.rn.f16.f32):
the bounds DESIGN.md section 4 and rover_slam_b200/csrc/common.cuh state for
  * the general split  v = hi + lo / 2^11                       (split_f32: every GEMM / convolution operand),
  * the attention probabilities  E = 2^11 p, P_hi = rn16(E), P_lo = rn16(E - P_hi)      (attn_kernel.cuh),
  * the attention values  V_hi = rn16(256 v), V_lo = rn16(256 v - V_hi)                   (RFE_ATTN_V_SCALE),
and that a split-fp16 dot product with three partial products reproduces the fp32 one to fp32 accuracy."""
import numpy as np

F16_MIN_NORMAL = 2.0 ** -14


def rn16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def split(v):
    hi = rn16(v)
    lo = rn16((np.asarray(v, np.float32) - hi) * np.float32(2048.0))
    return hi, lo


def test_general_split_is_22_bits():
    rng = np.random.RandomState(0)
    v = (rng.randn(200000) * np.exp(rng.uniform(-6, 4.8, 200000))).astype(np.float32)       # magnitudes down to denormal hi, up to ~400
    hi, lo = split(v)
    rec = hi.astype(np.float64) + lo.astype(np.float64) / 2048.0
    err = np.abs(rec - v.astype(np.float64))
    normal = np.abs(v) >= F16_MIN_NORMAL
    assert np.abs(v).max() < 65504 and normal.sum() > 190000
    assert (err[normal] <= 2.0 ** -22 * np.abs(v[normal])).all()                             # 22 significand bits while hi is a normal fp16
    assert (err[~normal] <= 2.0 ** -36 * 1.0001).all()                                       # below 6.1e-5: absolute (half a denormal step of lo / 2^11)
    assert (np.abs(lo[normal]) <= np.abs(hi[normal]) * 1.0001).all()                         # lo / 2^11 is at most half an ulp of hi


def test_attention_probability_planes():
    rng = np.random.RandomState(1)
    p = np.exp(-rng.uniform(0, 30, 300000)).astype(np.float32)                               # softmax numerators exp(s - max) in (9e-14, 1]
    p[:4] = [1.0, 0.5, 2.0 ** -14, 2.0 ** -30]
    E = p * np.float32(2048.0)
    hi = rn16(E)
    lo = rn16(E - hi)                                                                        # E - hi is exact in fp32 (the FHFMA)
    assert np.array_equal((E.astype(np.float64) - hi.astype(np.float64)).astype(np.float32), E - hi)
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / 2048.0
    err = np.abs(rec - p.astype(np.float64))
    # 21 bits relative while the low plane is a normal fp16 (p >= 2^-15 = 3e-5); below that the error is ABSOLUTE, half a denormal
    # step of the low plane = 2^-25 / 2^11 = 2^-36 of the row maximum (which is ~1): 2000 such keys perturb a row sum by < 3e-8
    assert (err <= np.maximum(2.0 ** -21 * p.astype(np.float64), 2.0 ** -36 * 1.0001)).all()
    big = p >= 2.0 ** -15
    assert big.sum() > 90000 and (err[big] / p[big] <= 2.0 ** -21).all()
    assert E.max() <= 2048.0 * 1.0 and 2048.0 * np.exp(3.4) < 65504                          # head-room: the hi-only max may be 3.4 too low


def test_attention_value_planes():
    rng = np.random.RandomState(2)
    v = (rng.randn(300000) * np.exp(rng.uniform(-12, 3.8, 300000))).astype(np.float32)       # up to ~ +-150 (largest LightGlue value seen: 46)
    v = v[np.abs(v) < 250]
    s = v * np.float32(256.0)
    hi = rn16(s)
    lo = rn16(s - hi)
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / 256.0
    err = np.abs(rec - v.astype(np.float64))
    assert np.isfinite(hi).all() and (err <= np.maximum(2.0 ** -21 * np.abs(v), 2.0 ** -25 / 256.0 * 1.0001)).all()


def test_three_product_dot_is_fp32_equivalent():
    """acc0 = a_hi.b_hi ; acc1 = a_hi.b_lo + a_lo.b_hi ; result acc0 + acc1 / 2^11 -- against the float64 dot product the error is
    of the order of fp32's own rounding (the dropped a_lo.b_lo term is 2^-22 relative)."""
    rng = np.random.RandomState(3)
    worst = 0.0
    for k in (64, 256, 512, 1152):
        a = (rng.randn(64, k) * 3).astype(np.float32)
        b = rng.randn(k, 32).astype(np.float32)
        ah, al = split(a)
        bh, bl = split(b)
        acc0 = ah.astype(np.float64) @ bh.astype(np.float64)
        acc1 = ah.astype(np.float64) @ bl.astype(np.float64) + al.astype(np.float64) @ bh.astype(np.float64)
        got = acc0 + acc1 / 2048.0
        ref = a.astype(np.float64) @ b.astype(np.float64)
        scale = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))                # sum of |terms|: the natural error scale
        worst = max(worst, (np.abs(got - ref) / scale).max())
        f32 = np.abs((a @ b).astype(np.float64) - ref) / scale
        assert (np.abs(got - ref) / scale).max() <= 2.0 ** -21
        assert (np.abs(got - ref) / scale).max() <= 4 * max(f32.max(), 2.0 ** -24)
    assert worst > 0


def test_lightglue_with_emulated_split_fp16_everywhere_matches_fp32():
    """End-to-end design check on the CPU: LightGlue with EVERY tensor-core contraction replaced by the arithmetic the kernels
    perform -- three-product split-fp16 linears, attention with the hi-only row maximum, E = 2^11 exp(s - max) probability planes and
    256-scaled value planes -- returns the fp32 oracle's match set with match scores within the stated tolerance."""
    import torch
    from oracle import lightglue_ref, synth
    from tests import parity

    def t16(x):                                   # fp32 tensor -> nearest fp16, back in float64
        return x.to(torch.float32).to(torch.float16).to(torch.float64)

    def split_t(x):
        x = x.to(torch.float32)
        hi = x.to(torch.float16).to(torch.float32)
        lo = ((x - hi) * 2048.0).to(torch.float16).to(torch.float64)
        return hi.to(torch.float64), lo

    def mm(name, x, wt):
        if name in ("posenc", "matchability"):    # CUDA-core fp32 kernels, not tensor-core contractions
            return x @ wt.t()
        xh, xl = split_t(x)
        wh, wl = split_t(wt)
        return (xh @ wh.t() + (xh @ wl.t() + xl @ wh.t()) / 2048.0).to(torch.float32)

    def attn(q, k, v):
        qs, ks = (q * lightglue_ref.ATTN_SCALE).to(torch.float32), (k * lightglue_ref.ATTN_SCALE).to(torch.float32)
        qh, ql = split_t(qs)
        kh, kl = split_t(ks)
        s_hh = qh @ kh.transpose(1, 2)
        s = (s_hh + (qh @ kl.transpose(1, 2) + ql @ kh.transpose(1, 2)) / 2048.0).to(torch.float32)
        mx = s_hh.to(torch.float32).max(-1, keepdim=True).values          # pass 1: hi*hi products only
        E = (2048.0 * torch.exp((s - mx).to(torch.float32))).to(torch.float32)
        ph = E.to(torch.float16).to(torch.float32)
        pl = t16(E - ph)
        vs = (v * 256.0).to(torch.float32)
        vh = vs.to(torch.float16).to(torch.float32)
        vl = t16(vs - vh)
        ph, vh = ph.to(torch.float64), vh.to(torch.float64)
        o = (ph @ vh + (ph @ vl + pl @ vh)) / (256.0 * E.to(torch.float64).sum(-1, keepdim=True))
        return o.to(torch.float32)

    k0, k1, d0, d1, perm = synth.lightglue_inputs(256, 456)
    kn0, kn1 = lightglue_ref.normalize_keypoints(k0, 480, 640), lightglue_ref.normalize_keypoints(k1, 480, 640)
    rm, rs = lightglue_ref.LightGlueRef()(kn0, kn1, d0, d1)
    em, es = lightglue_ref.LightGlueRef(mm=mm, attn=attn)(kn0, kn1, d0, d1)
    r = parity.compare_matches(rm.numpy(), rs.numpy(), em.numpy(), es.numpy())
    assert r["common"] == len(rm) == len(em) and len(rm) > 200 and r["mscore_maxabs"] < 2e-4, r      # measured: 254 / 254, 2.3e-5


def test_superpoint_with_emulated_split_fp16_convolutions_matches_fp32():
    """The same design check for the extractor: every convolution except conv1a (exact fp32 FMA on CUDA cores) computed as the three
    split-fp16 products, bias added in fp32 -- keypoint set identical to the fp32 oracle, scores and descriptors inside the tolerance."""
    import torch
    import torch.nn.functional as F
    from oracle import superpoint_ref, synth
    from tests import parity

    def split_t(x):
        x = x.to(torch.float32)
        hi = x.to(torch.float16).to(torch.float32)
        lo = ((x - hi) * 2048.0).to(torch.float16).to(torch.float64)
        return hi.to(torch.float64), lo

    def conv_fn(name, x, w, b, pad):
        if name == "conv1a":
            return F.conv2d(x, w, b, padding=pad)
        xh, xl = split_t(x)
        wh, wl = split_t(w)
        y = F.conv2d(xh, wh, None, padding=pad) + (F.conv2d(xh, wl, None, padding=pad) + F.conv2d(xl, wh, None, padding=pad)) / 2048.0
        return y.to(torch.float32) + b.reshape(1, -1, 1, 1)

    img = synth.frame(21, 120, 160)
    rk, rs, rd = superpoint_ref.SuperPointRef()(img)
    ek, es, ed = superpoint_ref.SuperPointRef(conv_fn=conv_fn)(img)
    r = parity.compare_keypoints(rk.numpy(), rs.numpy(), ek.numpy(), es.numpy())
    assert len(rk) > 50 and len(r["ref_idx"]) >= len(rk) - 1
    parity.compare_descriptors(rd.numpy()[r["ref_idx"]], ed.numpy()[r["tst_idx"]])


def test_gelu_erfc_formulation_is_fp32_equivalent():
    """ln_gelu_split_kernel computes GELU through erfc(|x|) = t (a1 + t (... a5)) exp(-x^2), t = 1 / (1 + p |x|)
    (Abramowitz & Stegun 7.1.26) instead of erff(): restated here in fp32 and compared with float64 -- it must stay inside
    the error envelope of the graph's own fp32 formulation 0.5 y (1 + erff(y / sqrt 2))."""
    import math
    import torch
    from scipy.special import erf as erf64
    f = np.float32
    y = np.concatenate([np.linspace(-12, 12, 400001), np.random.RandomState(0).randn(400000) * 2]).astype(f)
    ref = 0.5 * y.astype(np.float64) * (1 + erf64(y.astype(np.float64) / math.sqrt(2)))
    yt = torch.from_numpy(y)
    g_erff = ((yt * (torch.erf(yt / f(1.4142135381698608)) + 1)) * f(0.5)).numpy()          # lightglue_ref.py / the ONNX nodes
    x = np.abs(y) * f(0.70710678118654752)
    t = f(1) / (f(1) + f(0.3275911) * x)
    p = ((((f(1.061405429) * t + f(-1.453152027)) * t + f(1.421413741)) * t + f(-0.284496736)) * t + f(0.254829592)) * t
    c = (p * np.exp2((x * x * f(-1.4426950408889634)).astype(f)).astype(f)).astype(f)
    g = (f(0.5) * y * np.where(y < 0, c, f(2) - c)).astype(f)
    e_new, e_old = np.abs(g - ref).max(), np.abs(g_erff - ref).max()
    assert e_new <= 5e-7 and e_new <= 1.25 * e_old, (e_new, e_old)             # measured 4.2e-7 vs 4.4e-7
    assert np.abs(g - g_erff).max() <= 6e-7
