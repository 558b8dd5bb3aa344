#!/usr/bin/env python
"""Bring-up diagnostics on a real B200: per-stage max-abs errors of the CUDA path vs the CPU oracle.
Not a test (tests/ holds those); it never stops at the first mismatch so one GPU trip tells everything."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import lightglue_ref, superpoint_ref, synth  # noqa: E402
from rover_slam_b200 import FrontEnd  # noqa: E402


def section(name):
    print(f"\n=== {name} " + "=" * (70 - len(name)), flush=True)


def gemm_checks(fe):
    rng = np.random.RandomState(0)
    for (m, n, k) in [(128, 64, 64), (128, 128, 64), (128, 64, 256), (128, 128, 512), (300, 200, 512),
                      (77, 768, 256), (1000, 65, 256)]:
        a = rng.randn(m, k).astype(np.float32)
        b = rng.randn(n, k).astype(np.float32)
        bias = rng.randn(n).astype(np.float32)
        try:
            d = fe.debug_gemm(a, b, bias)
            ref = a.astype(np.float64) @ b.astype(np.float64).T + bias
            err = np.abs(d - ref).max()
            print(f"gemm {m}x{n}x{k}: maxabs err {err:.3e} (ref max {np.abs(ref).max():.2f})", flush=True)
            if err > 1e-3:
                # locate the error pattern
                bad = np.argwhere(np.abs(d - ref) > 1e-3)
                print("   first bad", bad[:5].tolist(), "count", len(bad), "d", d[tuple(bad[0])], "ref", ref[tuple(bad[0])])
                rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
                print("   bad rows", rows[:20].tolist(), "... bad cols", cols[:20].tolist())
        except Exception as e:
            print(f"gemm {m}x{n}x{k}: FAILED {e}", flush=True)


def sp_checks(fe, h, w, seeds):
    sp = superpoint_ref.SuperPointRef()
    imgs = np.stack([synth.frame(s, h, w) for s in seeds])
    t = time.time()
    feats = fe.extract(imgs)
    print(f"extract {imgs.shape}: {time.time() - t:.3f}s, counts {[len(f[0]) for f in feats]}", flush=True)
    B = len(seeds)
    for bi in range(B):
        taps = {}
        rk, rs, rd = sp(imgs[bi], taps)

        def nhwc(name, c, div):
            arr = fe.debug_read(name).reshape(B, h // div, w // div, c)[bi]
            return np.transpose(arr, (2, 0, 1))
        for name, key, c, div in [("sp.a1a", "relu1a", 64, 1), ("sp.pool1", "pool1", 64, 2), ("sp.pool2", "pool2", 64, 4),
                                  ("sp.pool3", "pool3", 128, 8), ("sp.feat", "feat", 128, 8), ("sp.dense", "dense_desc", 256, 8)]:
            try:
                got = nhwc(name, c, div)
                ref = taps[key][0].numpy()
                err = np.abs(got - ref)
                print(f"  img{bi} {name:9s} maxabs {err.max():.3e} (ref max {np.abs(ref).max():.2f}) mean {err.mean():.2e}", flush=True)
                if err.max() > 1e-2:
                    bad = np.argwhere(err > 1e-2)
                    print("     bad count", len(bad), "of", err.size, "first", bad[:3].tolist())
                    print("     bad channels", np.unique(bad[:, 0])[:16].tolist(), "rows", np.unique(bad[:, 1])[:16].tolist(),
                          "cols", np.unique(bad[:, 2])[:16].tolist())
            except Exception as e:
                print(f"  img{bi} {name}: FAILED {e}")
        heat = fe.debug_read("sp.heat").reshape(B, h, w)[bi]
        rh = taps["heatmap"][0].numpy()
        print(f"  img{bi} heat      maxabs {np.abs(heat - rh).max():.3e}", flush=True)
        nms = fe.debug_read("sp.nms").reshape(B, h, w)[bi]
        rn = taps["nms"][0].numpy()
        print(f"  img{bi} nms       maxabs {np.abs(nms - rn).max():.3e}  (mismatching px {(np.abs(nms - rn) > 1e-5).sum()})", flush=True)
        # NMS kernel exactness given OUR heat-map
        own = superpoint_ref.SuperPointRef.nms(torch.from_numpy(heat)[None])
        _, _, post = superpoint_ref.SuperPointRef.select(own)
        print(f"  img{bi} nms(own heat) mismatching px {(post[0].numpy() != nms).sum()}", flush=True)
        k, s, d = feats[bi]
        A = set(map(tuple, rk.numpy().tolist())); Bs = set(map(tuple, k.tolist()))
        print(f"  img{bi} keypoints ref {len(A)} ours {len(Bs)} common {len(A & Bs)}", flush=True)
        if len(A & Bs):
            idx_r = {tuple(p): i for i, p in enumerate(rk.numpy().tolist())}
            common = [(idx_r[tuple(p)], i) for i, p in enumerate(k.tolist()) if tuple(p) in idx_r]
            ir, it = np.array(common).T
            print(f"       scores maxabs {np.abs(rs.numpy()[ir] - s[it]).max():.3e}  desc maxabs {np.abs(rd.numpy()[ir] - d[it]).max():.3e}", flush=True)
    return feats


def lg_checks(fe, n, seed):
    lg = lightglue_ref.LightGlueRef()
    k0, k1, d0, d1, perm = synth.lightglue_inputs(n, seed)
    t = time.time()
    m, ms = fe.match(k0, k1, d0, d1, 480, 640)
    print(f"match n={n}: {time.time() - t:.3f}s -> {len(m)} matches", flush=True)
    taps = {}
    kn0 = lightglue_ref.normalize_keypoints(k0, 480, 640)
    kn1 = lightglue_ref.normalize_keypoints(k1, 480, 640)
    rm, rms = lg(kn0, kn1, d0, d1, taps)
    n0p = (n + 7) // 8 * 8
    x = fe.debug_read("lg.x").reshape(-1, 256)
    print(f"  x0 final maxabs {np.abs(x[:n] - taps['cross8.x0'].numpy()).max():.3e}  x1 {np.abs(x[n0p:n0p + n] - taps['cross8.x1'].numpy()).max():.3e}", flush=True)
    sim = fe.debug_read("lg.sim").reshape(n, -1)[:, :n]
    print(f"  sim maxabs {np.abs(sim - taps['sim'].numpy()).max():.3e}", flush=True)
    A = set(map(tuple, rm.numpy().tolist())); B = set(map(tuple, m.tolist()))
    print(f"  matches ref {len(A)} ours {len(B)} common {len(A & B)}", flush=True)
    if len(A & B) == len(A) == len(B):
        print(f"  mscores maxabs {np.abs(ms - rms.numpy()).max():.3e}", flush=True)


def main():
    torch.set_num_threads(os.cpu_count())
    section("create")
    fe = FrontEnd(max_batch=2, max_height=480, max_width=752, max_keypoints=4096)
    print("ctx ok", flush=True)
    for name, fn in [("gemm", lambda: gemm_checks(fe)),
                     ("superpoint 120x160", lambda: sp_checks(fe, 120, 160, [3, 4])),
                     ("lightglue n=256", lambda: lg_checks(fe, 256, 456)),
                     ("superpoint 480x640", lambda: sp_checks(fe, 480, 640, [0])),
                     ("lightglue n=1000", lambda: lg_checks(fe, 1000, 77))]:
        section(name)
        try:
            fn()
        except Exception:
            traceback.print_exc()
    print("launches", fe.kernel_launches())


if __name__ == "__main__":
    main()
