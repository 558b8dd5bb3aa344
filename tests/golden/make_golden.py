#!/usr/bin/env python
"""Generate the committed golden vectors by executing the REFERENCE's own ONNX graphs literally.

Runs only in the build container (needs /root/reference/onnxmodel/*.onnx).  The graphs are walked
node by node by oracle/onnx_interp.py (torch-CPU fp32, ONNX opset semantics) -- the stand-in for the
reference's ONNXRuntime-CPU session (superpoint_onnx.cc:133-136, lightglue_onnx.cpp:210-214), which
cannot be installed here (no onnxruntime wheel/so, no network).

    python tests/golden/make_golden.py

Outputs (small .npz files, committed):
  sp_640x480_seed0.npz        BASELINE config 1: image, keypoints, scores, every 16th descriptor, heat-map rows
  sp_752x480_seed100_{a,b}.npz  BASELINE config 3 frames (pair related by a (+12,+7) px shift)
  lg_752x480_seed100.npz      config 3 matches/mscores for the interpreter's own features of that pair
                              (+ fp16-exact copies of the inputs are NOT stored: inputs are re-derived, see test)
  lg_synth_n{256,512}.npz     config 4 style synthetic LightGlue inputs (seeded) -> matches, mscores
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import onnx_interp, synth, lightglue_ref  # noqa: E402

REF = os.environ.get("ROVER_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
DESC_STRIDE = 16
HEAT_STRIDE = 16


def run_sp(sp, img):
    x = torch.from_numpy(img.astype(np.float32) * np.float32(1.0 / 255.0))[None, None]
    o = sp.run({"image": x}, keep=["/Reshape_1_output_0", "/Div_output_0"])
    return (o["keypoints"][0].numpy(), o["scores"][0].numpy(), o["descriptors"][0].numpy(),
            o["/Reshape_1_output_0"][0].numpy(), o["/Div_output_0"][0].numpy())


def save_sp(name, img, k, s, d, heat, dense):
    np.savez_compressed(
        os.path.join(OUT, name), image=img, keypoints=k.astype(np.int16), scores=s,
        desc_rows=np.arange(0, len(k), DESC_STRIDE, dtype=np.int32), desc=d[::DESC_STRIDE],
        heat_rows=np.arange(0, heat.shape[0], HEAT_STRIDE, dtype=np.int32), heat=heat[::HEAT_STRIDE],
        dense_desc_px=dense[:, ::8, ::8].copy())


def main():
    torch.set_num_threads(os.cpu_count())
    sp = onnx_interp.Interpreter(os.path.join(REF, "onnxmodel", "superpoint.onnx"))
    lg = onnx_interp.Interpreter(os.path.join(REF, "onnxmodel", "lightglue_sim.onnx"))

    img = synth.frame(0, 480, 640)
    save_sp("sp_640x480_seed0.npz", img, *run_sp(sp, img))

    a, b = synth.frame_pair(100, 480, 752)
    ra, rb = run_sp(sp, a), run_sp(sp, b)
    save_sp("sp_752x480_seed100_a.npz", a, *ra)
    save_sp("sp_752x480_seed100_b.npz", b, *rb)
    kn0 = lightglue_ref.normalize_keypoints(ra[0], 480, 752)
    kn1 = lightglue_ref.normalize_keypoints(rb[0], 480, 752)
    o = lg.run({"kpts0": kn0[None], "kpts1": kn1[None], "desc0": ra[2][None], "desc1": rb[2][None]})
    np.savez_compressed(os.path.join(OUT, "lg_752x480_seed100.npz"),
                        matches=o["matches0"].numpy().astype(np.int32), mscores=o["mscores0"].numpy(),
                        n0=len(ra[0]), n1=len(rb[0]))

    for n in (256, 512):
        k0, k1, d0, d1, perm = synth.lightglue_inputs(n, 200 + n)
        kn0 = lightglue_ref.normalize_keypoints(k0, 480, 640)
        kn1 = lightglue_ref.normalize_keypoints(k1, 480, 640)
        o = lg.run({"kpts0": kn0[None], "kpts1": kn1[None], "desc0": d0[None], "desc1": d1[None]},
                   keep=["/log_assignment.8/Add_2_output_0"])
        S = o["/log_assignment.8/Add_2_output_0"][0].numpy()
        np.savez_compressed(os.path.join(OUT, f"lg_synth_n{n}.npz"),
                            matches=o["matches0"].numpy().astype(np.int32), mscores=o["mscores0"].numpy(),
                            S_diag=S[np.arange(n), np.argsort(perm)].copy(), S_row0=S[0].copy(),
                            input_sha=np.frombuffer(
                                __import__("hashlib").sha1(d1.tobytes() + k1.tobytes()).digest(), dtype=np.uint8))
        print(n, "matches", o["matches0"].shape[0])
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
