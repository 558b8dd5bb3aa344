#!/usr/bin/env python
"""bench.py -- frames/sec (SuperPoint extract + LightGlue match) on 640x480 synthetic frame pairs.

A "step" = one batch of PAIRS 640x480 frame pairs: both frames of every pair are extracted (one batched
SuperPoint launch sequence over 2*PAIRS frames), then every pair is matched.  frames/s = 2*PAIRS*steps / time.
  value : inputs already resident in HBM, features handed from the extractor to the matcher on the device
          (rfe_sp_extract_device + rfe_lg_match_slots), device-timed with CUDA events, max over ranks.
  e2e   : N = 1: the same work through the host API (rfe_pairs_submit / rfe_pairs_collect_begin_full / _end: pinned host images
          in; host keypoints, scores, fp32 DESCRIPTORS, matches and match scores out -- everything the reference's
          SPextractor::operator() + MatchingPoints_onnx hand back), every H2D / D2H copy inside the timed region, two batches in
          flight, descriptors returned on a copy stream.
          N > 1: BASELINE config 5 as stated -- the frame stream lives in rank 0's pinned host memory; per step rank 0 copies
          all ranks' frames to its GPU, NCCL scatters one block per rank, every rank extracts + matches, NCCL gathers the
          fixed-size match records to rank 0, which copies them to the host; ingest / gather run on a side stream under the
          compute (rover_slam_b200/shard.py, PairStream).  All of it inside the timed region.
  stream: the config-5 pipeline measured at every N (N = 1 included), so that its scaling can be read off one key.
  latency: N = 1 only: p50 / p99 of ONE SPextractor::operator() + ONE SPmatcher::MatchingPoints_onnx(Frame, Frame) at batch 1
          through the C++ class surface (rover_slam_b200/latency_driver), the reference's per-frame call pattern.
  --impl reference : the reference's CPU path for the same workload -- its two ONNX graphs restated on torch-CPU
          (oracle/, stand-in for ONNXRuntime-CPU which is not installable here), all host threads, one pair per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 480, 640
WORKLOAD = ("BASELINE config 5 shape: 640x480 synthetic frame pairs, SuperPoint on both frames + one LightGlue match per pair")
SP_FLOPS_PER_FRAME = 52.10e9
ATTN_DRAM_BYTES_PER_LAUNCH = 119_588_352                   # dram__bytes_read.sum + dram__bytes_write.sum of ONE attention launch at 8 pairs
                                                           # (profiles/r02_attn2_full.ncu-rep, final build: 103.18 MB + 16.41 MB; algorithmic bytes 131 MB)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1440.2), d.get("hbm_gbs", 6572.5), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """`nvidia-smi -lms 50` running beside the timed regions (one process, so the samples really fall under load)."""

    def __init__(self, index):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(index),
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) >= 7:
                self.rows.append(cells)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        rows = []
        for r in self.rows:
            try:
                rows.append((float(r[0]), float(r[1]), float(r[2]), r[3:7]))
            except ValueError:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        busy = [r for r in rows if r[2] > 0.5 * max(x[2] for x in rows)] or rows      # samples taken under load (by power draw)
        sm = sorted(r[0] for r in busy)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3][i].lower().startswith("active") for r in busy):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons, "samples": len(rows),
                "samples_under_load": len(busy), "power_w_max": max(r[2] for r in rows)}


def make_pairs(n_pairs, seed0):
    from oracle import synth   # input generator only (numpy); not the checker
    rng = np.random.RandomState(seed0)
    frames = np.empty((n_pairs, 2, H, W), np.uint8)
    for p in range(n_pairs):
        dx, dy = rng.randint(-16, 17, size=2)
        a, b = synth.frame_pair(seed0 * 1000 + 2 * p, H, W, shift=(int(dx), int(dy)))
        frames[p, 0], frames[p, 1] = a, b
    return frames


def pick_cpu_threads():
    """All the host threads the CPU path can USE: oversubscribing a shared 128-vCPU host makes torch-CPU slower,
    so time one small SuperPoint forward at a few thread counts and keep the fastest."""
    import torch
    from oracle import superpoint_ref, synth
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail})
    sp = superpoint_ref.SuperPointRef()
    img = synth.frame(3, 240, 320)
    best, best_t = cands[0], 1e30
    for c in cands:
        torch.set_num_threads(c)
        sp(img)
        t = time.perf_counter()
        sp(img)
        dt = time.perf_counter() - t
        if dt < best_t:
            best, best_t = c, dt
    return best, avail


def cpu_pair_seconds(frames_pair, threads):
    """Reference CPU path for one pair: 2 extracts + 1 match (torch-CPU restatement of the two ONNX graphs)."""
    import torch
    from oracle import lightglue_ref, superpoint_ref
    torch.set_num_threads(threads)
    sp = cpu_pair_seconds.sp = getattr(cpu_pair_seconds, "sp", None) or superpoint_ref.SuperPointRef()
    lg = cpu_pair_seconds.lg = getattr(cpu_pair_seconds, "lg", None) or lightglue_ref.LightGlueRef()
    t = time.perf_counter()
    with torch.no_grad():
        ka, _, da = sp(frames_pair[0])
        kb, _, db = sp(frames_pair[1])
        t1 = time.perf_counter()
        m, _ = lg(lightglue_ref.normalize_keypoints(ka.numpy(), H, W), lightglue_ref.normalize_keypoints(kb.numpy(), H, W), da, db)
    t2 = time.perf_counter()
    cpu_pair_seconds.last_split = ((t1 - t) / 2, t2 - t1)        # seconds per extraction, per match
    return t2 - t, len(m)


def batch1_latency(iters=40):
    """One SPextractor::operator() + one SPmatcher::MatchingPoints_onnx(Frame, Frame) per frame at batch 1 through the C++
    class surface (rover_slam_b200/latency_driver): p50 / p99 in milliseconds, host containers in and out."""
    import tempfile
    drv = os.path.join(ROOT, "rover_slam_b200", "latency_driver")
    if not os.path.exists(drv):
        return {"unavailable": "latency_driver not built (make host)"}
    frames = make_pairs(1, 11)
    with tempfile.TemporaryDirectory() as td:
        pa, pb = os.path.join(td, "a.raw"), os.path.join(td, "b.raw")
        frames[0, 0].tofile(pa)
        frames[0, 1].tofile(pb)
        env = dict(os.environ, ROVER_FE_WEIGHTS=os.path.join(ROOT, "weights", "rover_fe.rfw"))
        try:
            r = subprocess.run([drv, str(H), str(W), pa, pb, str(iters)], capture_output=True, text=True, env=env, timeout=300)
            out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        except Exception as e:          # noqa: BLE001
            return {"unavailable": f"latency_driver failed: {e!r}"}
    out["what"] = ("per frame: one SPextractor::operator() (640x480 u8 in; keypoints + N x 256 descriptors out) + one "
                   "SPmatcher::MatchingPoints_onnx(Frame, Frame) against the previous frame, batch 1, host containers, C++ class surface")
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads, avail = pick_cpu_threads()
    frames = make_pairs(1, 7)
    for _ in range(args.warmup):
        cpu_pair_seconds(frames[0], threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_pair_seconds(frames[0], threads)
        t += dt
    fps = 2 * args.steps / t
    line = {"metric": "frames/sec (extract+match) on 640x480", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": 1, "frames_per_step_per_gpu": 2,
                       "keypoints_per_frame": "about 2000 (no top-K, threshold 0.0005)",
                       "sample": "bounded sample of the workload: one pair of the same synthetic stream per step",
                       "reference_kind": "torch-CPU restatement of superpoint.onnx / lightglue_sim.onnx "
                       "(ONNXRuntime is not installable offline)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x 1 pair (2 extracts + 1 match); {threads} of {avail} host threads "
                                       "(fastest of a 5-point thread sweep)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=8, help="frame pairs per step per GPU")
    ap.add_argument("--cpu-pairs", type=int, default=2, help="pairs timed for cpu_baseline (rank 0, N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rover_slam_b200 import FrontEnd     # raises loudly when librover_fe.so is missing

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # SMs left to NCCL's kernels during the config-5 stream (rfe_set_sm_limit).  Measured at N = 2 with 0 / 4 / 8 SMs and NCCL
    # capped at 4 / 8 CTAs: stream / value = 0.963 / 0.964 / 0.947 -- the collectives do not disturb the persistent kernels, so 0.
    sm_reserve = int(os.environ.get("RFE_SM_RESERVE", "0")) if world > 1 else 0
    P = args.pairs
    B = 2 * P
    stream = torch.cuda.current_stream()
    fe = FrontEnd(device=local, stream=stream.cuda_stream, max_batch=B, max_height=H, max_width=W, max_keypoints=4096)

    # ---- input: rank 0 makes the synthetic stream, NCCL scatters one shard per rank (the path's only collective) ----
    from rover_slam_b200 import shard as sharding
    n_sets = 4                                  # distinct step inputs, cycled
    n_pairs_total = n_sets * P * world
    host_all = torch.from_numpy(make_pairs(n_pairs_total, 1)) if rank == 0 else None     # [world * n_sets * P, 2, H, W], rank r owns block r
    allf = host_all.to(dev) if rank == 0 else None
    block, valid = sharding.scatter_pairs(allf, n_pairs_total, (H, W), dev)     # [n_sets*P, 2, H, W] on this rank
    assert valid == n_sets * P
    shard = block.reshape(n_sets, B, H, W)
    del allf
    host_sets = shard.cpu().pin_memory()

    slots_a, slots_b = list(range(0, B, 2)), list(range(1, B, 2))

    def step_device(i):
        fe.extract_device(shard[i % n_sets].data_ptr(), H, W, W, B)
        fe.match_slots_batch(slots_a, slots_b, H, W, 0.0)

    h2d = d2h = 0

    host_np = [host_sets[i].numpy() for i in range(n_sets)]      # views of the pinned buffers

    def run_host(n_steps):
        # the public end-to-end calls: pinned host frames in (rfe_pairs_submit), host keypoints + matches out
        # (rfe_pairs_collect), two batches in flight so that the host-side matcher set-up of batch i overlaps SuperPoint
        # of batch i+1 on the GPU.  Every step copies its own inputs H2D and its own results D2H.
        fe.pairs_submit(host_np[0])
        if n_steps > 1:
            fe.pairs_submit(host_np[1 % n_sets])
        for i in range(n_steps):
            fe.pairs_collect_begin(want_desc=True)      # enqueue LightGlue + result copies of batch i (descriptors on the copy stream)
            if i + 2 < n_steps:
                fe.pairs_submit(host_np[(i + 2) % n_sets])   # queue the next batch behind it BEFORE blocking
            kpts, res, feats = fe.pairs_collect_end()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # keypoint counts of every input set (for the attention kernel's algorithmic FLOPs): one untimed pass
    attn_flops_sets = []
    for i in range(n_sets):
        step_device(i)
        fe.sync()
        n = [len(fe.read_slot(b, want_desc=False)[0]) for b in range(B)]
        # per attention launch: problems (n0,n0), (n1,n1) [self] or (n0,n1), (n1,n0) [cross]; 4 heads x (QK^T + PV) x 2*nq*nk*64
        self_f = sum(1024.0 * n[2 * p] ** 2 + 1024.0 * n[2 * p + 1] ** 2 for p in range(P))
        cross_f = sum(2 * 1024.0 * n[2 * p] * n[2 * p + 1] for p in range(P))
        attn_flops_sets.append((self_f + cross_f) / 2.0)          # mean over the self and the cross launch of a layer

    # ---- device-resident throughput ("value") ----
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    fe.profile(True, select="lg.attn")          # events only around the dominant kernel (18 launches per step)
    launches0 = fe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(i)
    e1.record(stream)
    barrier()
    launches = fe.kernel_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    attn_ms, attn_n = fe.profile_read("lg.attn", reset=True)
    fe.profile(False)
    nmatch = [len(fe.read_result(p)[0]) for p in range(P)]
    frames_total = 2 * P * args.steps * world
    value = frames_total / (ms_total / 1e3)
    attn_flops = sum(attn_flops_sets[i % n_sets] for i in range(args.steps)) / args.steps     # mean per launch

    e_steps = max(3, args.steps)
    # ---- BASELINE config 5: rank-0 host stream -> scatter -> extract + match -> gather -> rank-0 host ("stream") ----
    cap = fe.cap
    words = P * cap * 3 + P                              # per rank and step: matches [P][cap][2] i32, mscores [P][cap] f32, counts [P]

    def stream_compute(in_block, out_record, step):
        fe.extract_device(in_block.data_ptr(), H, W, W, B)
        fe.match_slots_batch(slots_a, slots_b, H, W, 0.0)
        base = out_record.data_ptr()
        fe.copy_results_device(P, base, base + 4 * P * cap * 2, base + 4 * P * cap * 3)

    ps = sharding.PairStream((B, H, W), words, dev, stream_compute)
    n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fe.set_sm_limit(n_sms - sm_reserve if sm_reserve else 0)
    stream_sets = None
    if rank == 0:                                       # [n_sets][world][B][H][W] in pinned host memory
        stream_sets = host_all.reshape(world, n_sets, B, H, W).transpose(0, 1).contiguous().pin_memory()
    last_counts = {}

    def consume(i, rec):
        last_counts["c"] = rec[:, P * cap * 3:].clone()

    ps.run(2, (lambda i: stream_sets[i % n_sets]) if rank == 0 else None)
    barrier()
    b0 = (ps.h2d_bytes, ps.d2h_bytes, ps.collective_bytes)
    t0 = time.perf_counter()
    ps.run(e_steps, (lambda i: stream_sets[i % n_sets]) if rank == 0 else None, consume)
    barrier()
    stream_s = max_over_ranks(time.perf_counter() - t0)
    stream_fps = 2 * P * e_steps * world / stream_s
    s_h2d, s_d2h, s_coll = ((ps.h2d_bytes - b0[0]) // e_steps, (ps.d2h_bytes - b0[1]) // e_steps, (ps.collective_bytes - b0[2]) // e_steps)
    fe.set_sm_limit(0)

    # ---- end to end through the host API ("e2e" at N = 1) ----
    run_host(2)
    barrier()
    tb0 = fe.transfer_bytes()
    t0 = time.perf_counter()
    run_host(e_steps)
    barrier()
    tb1 = fe.transfer_bytes()
    h2d, d2h = tb1[0] - tb0[0], tb1[1] - tb0[1]          # bytes the library actually copied (counted at the cudaMemcpyAsync calls)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = 2 * P * e_steps * world / e2e_s
    sampler.stop()

    # ---- per-kernel breakdown (untimed extra pass, every launch bracketed by events) ----
    fe.profile(True)
    for i in range(3):
        step_device(i)
    fe.sync()
    all_ms, all_n = fe.profile_read(None)
    breakdown = {}
    for tag in ["sp.conv1a", "sp.conv1b", "sp.conv2", "sp.conv3", "sp.conv4", "sp.convPa", "sp.convDa", "sp.convPb", "sp.convDb", "sp.nms",
                "sp.select", "sp.desc_sample", "lg.prepare", "lg.wqkv", "lg.attn", "lg.out_proj", "lg.ffn0", "lg.ln_gelu", "lg.ffn3",
                "lg.to_qk", "lg.to_v", "lg.to_out", "lg.final_proj", "lg.matchability", "lg.sim", "lg.assign"]:
        ms, n = fe.profile_read(tag)
        if n:
            breakdown[tag] = round(1e3 * ms / 3, 1)
    attn_share = breakdown.get("lg.attn", 0.0) / (1e3 * all_ms / 3) if all_ms else None
    fe.profile_read(None, reset=True)
    fe.profile(False)

    # ---- labelled fast mode (rfe_set_fast_mode: hi-only fp16 MMAs in SuperPoint's 3x3 convolutions), N = 1 only.  NOT the
    # parity path and not part of any number above: the same device-resident step, timed again, with the set overlap of its
    # keypoints and matches against the exact path on the same frames ----
    fast_block = None
    if world == 1:
        def snapshot():
            step_device(0)
            fe.sync()
            kps = [fe.read_slot(b, want_desc=False)[0] for b in range(B)]
            mts = [fe.read_result(p)[0] for p in range(P)]
            return kps, mts

        def coords(k):
            return set(map(tuple, np.asarray(k).tolist()))
        ek, em = snapshot()
        fe.set_fast_mode(True)
        try:
            for i in range(3):
                step_device(i)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record(stream)
            for i in range(args.steps):
                step_device(i)
            f1.record(stream)
            torch.cuda.synchronize()
            fast_ms = f0.elapsed_time(f1) / args.steps
            fk, fm = snapshot()
        finally:
            fe.set_fast_mode(False)
        kp_ov = [len(coords(a) & coords(b)) / max(len(a), 1) for a, b in zip(ek, fk)]
        m_ov = []
        for p_ in range(P):
            e = {(tuple(ek[2 * p_][i]), tuple(ek[2 * p_ + 1][j])) for i, j in em[p_].tolist()}
            f = {(tuple(fk[2 * p_][i]), tuple(fk[2 * p_ + 1][j])) for i, j in fm[p_].tolist()}
            m_ov.append(len(e & f) / max(len(e), 1))
        fast_block = {"value": 2 * P / (fast_ms / 1e3), "unit": "frames/s", "ms_per_step": fast_ms, "steps": args.steps,
                      "keypoint_overlap_mean": float(np.mean(kp_ov)), "keypoint_overlap_min": float(np.min(kp_ov)),
                      "match_overlap_mean": float(np.mean(m_ov)), "match_overlap_min": float(np.min(m_ov)),
                      "what": "NOT the parity path, not the default, not `value`: hi-only fp16 MMAs (1 instead of 3 per MAC) in SuperPoint's "
                              "3x3 convolutions, LightGlue unchanged; overlap = share of the exact path's keypoints (by pixel) / matches "
                              "(by pixel pair) that the fast path also returns, same frames"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tf_peak, hbm_peak, peak_src = peaks()
    per_launch_ms = attn_ms / max(attn_n, 1)
    achieved = attn_flops / (per_launch_ms / 1e3) / 1e12 if attn_n else 0.0
    line = {
        "metric": "frames/sec (extract+match) on 640x480", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16x2-split (fp32-equivalent), f32 accumulate", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "pairs_per_step_per_gpu": P, "frames_per_step_per_gpu": B, "keypoints_per_frame": "about 2000 (no top-K, threshold 0.0005)",
                   "parallelism": f"dp{world} (independent pairs per GPU; NCCL only scatters the input shards)",
                   "l2": f"per-step activation working set about {0.31 * B:.1f} GB >> 126 MB L2; 4 distinct input sets cycled",
                   "matches_last_step": nmatch},
        "clocks": sampler.summary(),
        # N = 1: the reference-facing host API (keypoints, scores, descriptors, matches back on the host);
        # N > 1: BASELINE config 5, rank-0 host stream -> NCCL scatter -> compute -> NCCL gather -> rank-0 host
        "e2e": ({"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d // e_steps, "d2h_bytes_per_step": d2h // e_steps,
                 "steps": e_steps, "path": "host API: rfe_pairs_submit / rfe_pairs_collect_begin_full / _end, pinned buffers, "
                 "returns keypoints + scores + fp32 descriptors + matches + match scores"} if world == 1 else
                {"value": stream_fps, "unit": "frames/s", "h2d_bytes_per_step": s_h2d, "d2h_bytes_per_step": s_d2h, "steps": e_steps,
                 "path": "config 5 stream: rank-0 pinned host frames -> H2D -> NCCL scatter -> extract + match on every rank -> "
                 "NCCL gather of match records -> D2H on rank 0 (independent pairs: both frames of every pair extracted)"}),
        "stream": {"value": stream_fps, "unit": "frames/s", "steps": e_steps, "h2d_bytes_per_step": s_h2d, "d2h_bytes_per_step": s_d2h,
                   "collective": "NCCL scatter (u8 frames) + gather (fixed-size match records), side stream, double buffered" if world > 1 else "none (one rank)",
                   "collective_bytes_per_step": s_coll, "host_api_e2e": e2e,
                   "sms_left_to_nccl": sm_reserve, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS") if world > 1 else None,
                   "match_counts_last_step": last_counts.get("c").tolist() if last_counts.get("c") is not None else None,
                   "variant": "independent pairs (2 extracts + 1 match per pair); the SLAM-shaped stream (pair p = frames p, p+1: one "
                   "extract + one match per frame) needs half the extractions per pair and is not what this number measures"},
        "gpu_launches": launches,
        "fast_mode": fast_block,
        "roofline": {"bound": "tensor", "kernel": "attn2_kernel (lg.attn_self / lg.attn_cross: persistent fused 4-head attention of all pairs, 18 launches per step)",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak if tf_peak else None,
                     "traffic": ATTN_DRAM_BYTES_PER_LAUNCH if P == 8 else None, "peak_source": peak_src, "launch_ms": per_launch_ms, "launches_timed": attn_n,
                     "algorithmic_flops_per_launch": attn_flops,
                     "executed_mma_flops_per_launch": 3.5 * attn_flops,
                     "share_of_step": attn_share,
                     "note": "split-fp16: 3 MMAs per algorithmic MAC for QK^T and PV plus a hi-only max pass = 3.5x: frac <= 0.286 by construction; "
                             "traffic = DRAM bytes of one launch from the ncu --set full capture profiles/r02_attn2_full.ncu-rep (8 pairs per step)"},
        "kernel_us_per_step": breakdown,
    }
    if world == 1:
        line["latency"] = batch1_latency()
    if world == 1 and args.cpu_pairs > 0:
        threads, avail = pick_cpu_threads()
        cpu_frames = make_pairs(args.cpu_pairs, 7)
        cpu_pair_seconds(cpu_frames[0], threads)          # warm-up
        tt = 0.0
        for p in range(args.cpu_pairs):
            dt, _ = cpu_pair_seconds(cpu_frames[p], threads)
            tt += dt
        if isinstance(line.get("latency"), dict):
            ex_s, ma_s = cpu_pair_seconds.last_split
            line["latency"]["cpu_frame_ms"] = round(1e3 * (ex_s + ma_s), 1)      # the CPU arm's time for the same 1 extract + 1 match
        line["cpu_baseline"] = {"value": 2 * args.cpu_pairs / tt, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_pairs} pairs (2 extracts + 1 match each) of the same synthetic stream, after 1 warm-up pair; "
                                          f"{threads} of {avail} host threads (fastest of a thread sweep)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
