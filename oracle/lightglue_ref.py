"""CPU restatement of the reference's LightGlue graph -- TEST INFRASTRUCTURE ONLY.

Oracle for the matcher half of the hot path: what ONNXRuntime executes for
/root/reference/onnxmodel/lightglue_sim.onnx inside
`LightGlueDecoupleOnnxRunner::Matcher_Inference` (src/Matchers/lightglue_onnx.cpp:162-240, Run at
:210-214), preceded by `Matcher_PreProcess`/`NormalizeKeypoints` (lightglue_onnx.cpp:140-159,
src/Matchers/transform.cpp:19-32) and followed by `Matcher_PostProcess_fused`
(lightglue_onnx.cpp:396-482, the score>thresh scatter at :437-453).
Node numbers are positions in `graph.node`.

Parity pin: no reference test or golden vector exists for this path ("parity unpinned" by the
reference); this restatement is pinned against a literal execution of the reference's ONNX graph
(oracle/onnx_interp.py, tests/test_oracle.py) and the golden vectors it produced (tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import weights as _weights

N_LAYERS = 9
HEADS = 4
HEAD_DIM = 64
FILTER_THRESHOLD = 0.1          # /Constant_6 (node 1515)
ATTN_SCALE = 0.3535533845424652  # 64**-0.25, applied to q and to k (nodes 50-51)


def normalize_keypoints(kpts_px: np.ndarray, h: int, w: int) -> np.ndarray:
    """transform.cpp:19-32: (kpt - (w/2, h/2)) / (max(w,h)/2) in fp32."""
    shift = np.array([np.float32(w) / np.float32(2), np.float32(h) / np.float32(2)], dtype=np.float32)
    scale = np.float32(max(w, h)) / np.float32(2)
    return ((kpts_px.astype(np.float32) - shift) / scale).astype(np.float32)


class LightGlueRef:
    def __init__(self, blob: dict | None = None, dtype=torch.float32, mm=None, attn=None):
        blob = blob if blob is not None else _weights.load()
        self.dtype = dtype
        self.p = {k[3:]: torch.from_numpy(v).to(dtype) for k, v in blob.items() if k.startswith("lg.")}
        # blob stores [out,in]; x @ W^T == ONNX MatMul(x, W_onnx)
        self.mm = mm or (lambda name, x, wt: x @ wt.t())
        # softmax(q k^T / 8) v per head, q and k each scaled by 64^-1/4 first (nodes 50-54); hook: numerics experiments
        self.attn = attn or (lambda q, k, v: torch.softmax((q * ATTN_SCALE) @ (k.transpose(1, 2) * ATTN_SCALE), -1) @ v)

    def lin(self, name, x):
        y = self.mm(name, x, self.p[name + ".w"])
        if name + ".b" in self.p:
            y = self.p[name + ".b"] + y
        return y

    def posenc(self, k):
        """nodes 0-17.  k [N,2] normalised -> e [2,N,64] (cos / sin, each frequency repeated twice)."""
        p = self.mm("posenc", k, self.p["posenc.w"])                 # [N,32]
        e = torch.stack([torch.cos(p), torch.sin(p)], 0).unsqueeze(-1)   # [2,N,32,1]
        return torch.cat([e, e], -1).reshape(2, -1, 64)

    @staticmethod
    def rot_half(t):
        t2 = t.reshape(*t.shape[:-1], 32, 2)
        return torch.stack([-t2[..., 1], t2[..., 0]], -1).reshape(t.shape)

    def rope(self, e, t):
        return t * e[0] + self.rot_half(t) * e[1]

    def ffn(self, pre, x, m):
        h = self.lin(pre + ".ffn0", torch.cat([x, m], -1))
        h = F.layer_norm(h, (512,), self.p[pre + ".ln.w"], self.p[pre + ".ln.b"], eps=1e-5)
        h = (h * (torch.erf(h / 1.4142135381698608) + 1.0)) * 0.5
        return x + self.lin(pre + ".ffn3", h)

    def self_block(self, i, x, e):
        n = x.shape[0]
        qkv = self.lin(f"l{i}.self.wqkv", x).reshape(n, HEADS, HEAD_DIM, 3).permute(1, 0, 2, 3)   # [4,N,64,3]
        q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
        q, k = self.rope(e, q), self.rope(e, k)
        a = self.attn(q, k, v)                                                                   # [4,N,64]
        m = self.lin(f"l{i}.self.out_proj", a.permute(1, 0, 2).reshape(n, 256))
        return self.ffn(f"l{i}.self", x, m)

    def cross_block(self, i, x0, x1):
        def heads(t):
            return t.reshape(t.shape[0], HEADS, HEAD_DIM).permute(1, 0, 2)
        pre = f"l{i}.cross"
        qk0, qk1 = heads(self.lin(pre + ".to_qk", x0)), heads(self.lin(pre + ".to_qk", x1))
        v0, v1 = heads(self.lin(pre + ".to_v", x0)), heads(self.lin(pre + ".to_v", x1))
        m0 = self.attn(qk0, qk1, v1)
        m1 = self.attn(qk1, qk0, v0)
        m0 = self.lin(pre + ".to_out", m0.permute(1, 0, 2).reshape(-1, 256))
        m1 = self.lin(pre + ".to_out", m1.permute(1, 0, 2).reshape(-1, 256))
        return self.ffn(pre, x0, m0), self.ffn(pre, x1, m1)

    def log_assignment(self, x0, x1):
        """nodes 1480-1501 -> S [N0,N1]."""
        md0 = self.lin("final_proj", x0) / 4.0
        md1 = self.lin("final_proj", x1) / 4.0
        sim = self.mm("sim", md0, md1)                                # md0 @ md1^T
        z0 = self.lin("matchability", x0)                             # [N0,1]
        z1 = self.lin("matchability", x1)
        cert = torch.log(torch.sigmoid(z0)) + torch.log(torch.sigmoid(z1)).t()
        return (torch.log_softmax(sim, 1) + torch.log_softmax(sim, 0)) + cert, sim

    @staticmethod
    def filter_matches(S):
        """nodes 1502-1525 -> matches i64 [K,2], mscores [K]."""
        max0, m0 = S.max(dim=1)
        _, m1 = S.max(dim=0)
        mutual0 = torch.arange(S.shape[0]) == m1[m0]
        ms0 = torch.where(mutual0, torch.exp(max0), torch.zeros_like(max0))
        idx = torch.nonzero(ms0 > FILTER_THRESHOLD)[:, 0]
        return torch.stack([idx, m0[idx]], -1), ms0[idx]

    def __call__(self, kn0, kn1, d0, d1, taps: dict | None = None):
        """kn*: normalised keypoints [N,2] f32; d*: [N,256] f32.  Returns (matches i64 [K,2], mscores f32 [K])."""
        kn0, kn1 = (torch.as_tensor(a).to(self.dtype) for a in (kn0, kn1))
        x0, x1 = (torch.as_tensor(a).to(self.dtype) for a in (d0, d1))
        if x0.shape[0] == 0 or x1.shape[0] == 0:
            return torch.zeros(0, 2, dtype=torch.int64), torch.zeros(0, dtype=self.dtype)
        e0, e1 = self.posenc(kn0), self.posenc(kn1)
        for i in range(N_LAYERS):
            x0 = self.self_block(i, x0, e0)
            x1 = self.self_block(i, x1, e1)
            if taps is not None:
                taps[f"self{i}.x0"], taps[f"self{i}.x1"] = x0, x1
            x0, x1 = self.cross_block(i, x0, x1)
            if taps is not None:
                taps[f"cross{i}.x0"], taps[f"cross{i}.x1"] = x0, x1
        S, sim = self.log_assignment(x0, x1)
        if taps is not None:
            taps["sim"], taps["S"] = sim, S
        return self.filter_matches(S)


def scatter_matches(matches, mscores, n0: int, thresh: float, vn=None):
    """lightglue_onnx.cpp:437-453: vnMatches12[i] = j for mscore > thresh.  Returns (vn, count)."""
    vn = np.full(n0, -1, dtype=np.int32) if vn is None else vn
    cnt = 0
    for (i, j), s in zip(np.asarray(matches).tolist(), np.asarray(mscores).tolist()):
        if s > thresh:
            vn[i] = j
            cnt += 1
    return vn, cnt
