"""CPU restatement of the reference's SuperPoint graph -- TEST INFRASTRUCTURE ONLY.

This is the oracle for the extractor half of the hot path.  It restates, in readable torch-CPU
fp32 (or fp64 for calibration), what ONNXRuntime executes for
/root/reference/onnxmodel/superpoint.onnx when the reference calls
`SuperPointOnnxRunner::Extractor_Inference` (src/Extractors/superpoint_onnx.cc:88-162, Run at :133-136)
after `NormalizeImage` (src/Matchers/transform.cpp:3-17: u8 -> f32 * 1/255).
Graph node numbers in the comments are positions in `graph.node`.

Parity pin: the reference ships no test or golden vector for this path ("parity unpinned" by the
reference itself); this restatement is pinned instead against a literal execution of the reference's
own ONNX graph (oracle/onnx_interp.py) -- tests/test_oracle.py -- and against the committed
golden vectors in tests/golden/ that were produced by that literal execution.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import weights as _weights

CONVS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
         "convPa", "convPb", "convDa", "convDb"]
NMS_RADIUS = 4          # MaxPool kernel 9, pads 4 (nodes 59-84)
BORDER = 4              # ScatterND x4 (nodes 85-359)
THRESHOLD = 0.0005      # /Constant_116 (node 360-363)


class SuperPointRef:
    def __init__(self, blob: dict | None = None, dtype=torch.float32, conv_fn=None):
        blob = blob if blob is not None else _weights.load()
        self.dtype = dtype
        self.w, self.b = {}, {}
        for c in CONVS:
            # blob holds OHWI; torch conv2d wants OIHW
            self.w[c] = torch.from_numpy(blob[f"sp.{c}.w"]).permute(0, 3, 1, 2).contiguous().to(dtype)
            self.b[c] = torch.from_numpy(blob[f"sp.{c}.b"]).to(dtype)
        self.conv_fn = conv_fn or (lambda name, x, w, b, pad: F.conv2d(x, w, b, padding=pad))

    def _conv(self, name, x, pad):
        return self.conv_fn(name, x, self.w[name], self.b[name], pad)

    # ---- dense part: nodes 1-57 and 399-413 ---------------------------------------------
    def backbone(self, x, taps: dict | None = None):
        """x: [B,1,H,W] in [0,1].  Returns (heatmap [B,H,W], dense descriptors [B,256,H/8,W/8])."""
        def tap(k, v):
            if taps is not None:
                taps[k] = v
        x = torch.relu(self._conv("conv1a", x, 1)); tap("relu1a", x)
        x = torch.relu(self._conv("conv1b", x, 1))
        x = F.max_pool2d(x, 2, 2); tap("pool1", x)                       # /pool/MaxPool_output_0
        x = torch.relu(self._conv("conv2a", x, 1))
        x = torch.relu(self._conv("conv2b", x, 1))
        x = F.max_pool2d(x, 2, 2); tap("pool2", x)
        x = torch.relu(self._conv("conv3a", x, 1))
        x = torch.relu(self._conv("conv3b", x, 1))
        x = F.max_pool2d(x, 2, 2); tap("pool3", x)
        x = torch.relu(self._conv("conv4a", x, 1))
        f = torch.relu(self._conv("conv4b", x, 1)); tap("feat", f)        # /relu_7/Relu_output_0
        # detector head, nodes 20-57
        s = torch.relu(self._conv("convPa", f, 1))
        s = self._conv("convPb", s, 0); tap("logits", s)                 # [B,65,h,w]
        s = torch.softmax(s, 1)[:, :64]
        b, _, h, w = s.shape
        s = s.permute(0, 2, 3, 1).reshape(b, h, w, 8, 8).permute(0, 1, 3, 2, 4).reshape(b, h * 8, w * 8)
        tap("heatmap", s)
        # descriptor head, nodes 399-413
        d = torch.relu(self._conv("convDa", f, 1))
        d = self._conv("convDb", d, 0)
        nrm = d.abs().pow(2.0).sum(1, keepdim=True).pow(0.5).clamp(min=1e-12)
        d = d / nrm; tap("dense_desc", d)
        return s, d

    # ---- NMS + border + threshold: nodes 58-398 -----------------------------------------
    @staticmethod
    def nms(scores):
        """scores [B,H,W] -> scores with non-maxima zeroed (2 suppression iterations, 9x9)."""
        k, p = 2 * NMS_RADIUS + 1, NMS_RADIUS

        def mp(t):
            return F.max_pool2d(t[:, None], k, 1, p)[:, 0]            # -inf padding
        zeros = torch.zeros_like(scores)
        max_mask = scores == mp(scores)
        for _ in range(2):
            supp_mask = mp(max_mask.to(scores.dtype)) > 0
            supp_scores = torch.where(supp_mask, zeros, scores)
            new_max = supp_scores == mp(supp_scores)
            max_mask = max_mask | (new_max & (~supp_mask))
        return torch.where(max_mask, scores, zeros)

    @staticmethod
    def select(s):
        """s [1,H,W] post-NMS -> (keypoints i64 [N,2] as (x,y), scores [N]) in row-major order."""
        s = s.clone()
        s[:, :BORDER, :] = -1
        s[:, :, :BORDER] = -1
        s[:, -BORDER:, :] = -1
        s[:, :, -BORDER:] = -1
        idx = torch.nonzero(s[0] > THRESHOLD)                           # rows (y,x), row-major
        sc = s[0][idx[:, 0], idx[:, 1]]
        return idx.flip(1).contiguous(), sc, s

    # ---- descriptor sampling: nodes 414-485 ---------------------------------------------
    @staticmethod
    def sample(dense, kpts):
        """dense [1,256,h,w] (L2-normalised), kpts i64 [N,2] (x,y) -> [N,256] unit descriptors."""
        _, c, h, w = dense.shape
        k = kpts.to(dense.dtype) - 4.0 + 0.5
        gx = k[:, 0] / (float(w * 8) - 4.0 - 0.5)
        gy = k[:, 1] / (float(h * 8) - 4.0 - 0.5)
        g = torch.stack([gx, gy], -1) * 2.0 - 1.0
        dd = F.grid_sample(dense, g.view(1, 1, -1, 2), mode="bilinear", padding_mode="zeros",
                           align_corners=True)
        dd = dd.reshape(1, c, -1)
        nrm = dd.abs().pow(2.0).sum(1, keepdim=True).pow(0.5).clamp(min=1e-12)
        return (dd / nrm)[0].t().contiguous()

    # ---- whole graph ----------------------------------------------------------------------
    def __call__(self, image_u8: np.ndarray, taps: dict | None = None):
        """image_u8: [H,W] uint8.  Returns (kpts i64 [N,2] xy, scores f32 [N], desc f32 [N,256])."""
        x = torch.from_numpy(image_u8.astype(np.float32) * np.float32(1.0 / 255.0))   # transform.cpp:8
        x = x.to(self.dtype)[None, None]
        heat, dense = self.backbone(x, taps)
        nmsed = self.nms(heat)
        kpts, sc, post = self.select(nmsed)
        if taps is not None:
            taps["nms"] = post
        desc = self.sample(dense, kpts)
        return kpts, sc, desc
