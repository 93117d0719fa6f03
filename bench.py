#!/usr/bin/env python
"""bench.py -- LiDAR frames/s of the chessboard-corner hot path on synthetic 16x1800 XYZI clouds.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU path on the box's host cores)

A "step" is one pass of the whole hot path (K1 roi_filter .. K5 board_fit, + the exchange of the corners between ranks
when N>1: fused into K5 + one barrier, or an NCCL all-gather) over one batch of --frames synthetic VLP-16 frames per
GPU (BASELINE.json configs[2]; 8 ranks x 1024
frames = configs[4]).  Prints ONE JSON line (rank 0).
  value        frames/s with the batch already resident in HBM (CUDA events around exactly K steps, max over ranks)
  e2e          same metric through the C ABI with HOST buffers: pinned H2D of the batch + D2H of corners/status
               inside the timed region
  roofline     dominant kernel of the step vs the measured HBM peak (algorithmic bytes = SURVEY.md 8(d))
  cpu_baseline the CPU oracle (a restatement; the reference needs ROS/PCL/Ceres and cannot be built here) timed
               on a bounded sample in the same run
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lidar_frames_per_sec_16x1800_xyzi"
UNIT = "frames/s"
BYTES_PER_FRAME = 28800 * 16 + 12 + 35 * 12 + 4  # 461,236 B algorithmic bytes per VLP-16 frame
HBM_FALLBACK_GBS = 6650.0                        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel, from the newest committed
    `ncu --set full` capture of this workload (1024 frames): profiles/ncu_traffic_*.json, written by
    tools/ncu_summary.py next to the text summaries. Returns {stage: bytes}."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_traffic_*.json")))
    if not paths:
        return {}
    try:
        return {k: v["dram_bytes"] for k, v in json.load(open(paths[-1])).items()}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def make_workload(n_frames: int, start: int, want_truth: bool = False):
    from lidar_camera_calibration_b200 import synth
    t = time.time()
    if want_truth:
        pts, clicks, gt = synth.make_batch(n_frames, start=start, want_truth=True)
        return pts, clicks, gt, time.time() - t
    pts, clicks = synth.make_batch(n_frames, start=start)
    return pts, clicks, time.time() - t


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra @ Rb.T) - 1.0) / 2.0, -1.0, 1.0)))


def pnp_leg(api, ldist, synth, torch, dev, rank, world, start, gt, corners_all, status_all, reps=3):
    """BASELINE.json configs[4], the step after the all-gather: order the corners of ALL frames, one Huber PnP over
    all pairs on rank 0's GPU (csrc/ilcc2_pnp.cu), against 2-D corners projected through a planted extrinsic.
    Outside the timed frames/s region; timed on its own with CUDA events."""
    uv = torch.from_numpy(synth.project_truth(gt, start=start)).to(dev)
    uv_all = ldist.allgather_rows(uv, world)
    if rank != 0:
        return None
    pnp = api.PnpSolver(dev.index)
    Tax = api.axis_roughly("pointgrey")
    bw, bh = synth.SCENE_VLP16.board.corners_long, synth.SCENE_VLP16.board.corners_short
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ms_order = ms_solve = 0.0
    for i in range(reps + 1):
        e0.record()
        p3, p2, valid = pnp.order_corners(corners_all, Tax, bw, bh, status=status_all, cam_corners=uv_all)
        e1.record()
        r, t, sm = pnp.solve(p3, p2, synth.PLANTED_CAMERA, valid=valid)
        e2.record()
        torch.cuda.synchronize(dev)
        if i > 0:
            ms_order += e0.elapsed_time(e1) / reps
            ms_solve += e1.elapsed_time(e2) / reps
    T = api.compose_extrinsic(r, t, Tax)
    G = synth.PLANTED_EXTRINSIC
    return {"pairs": int(valid.sum().item()), "frames": int(status_all.shape[0]), "order_ms": ms_order,
            "solve_ms": ms_solve, "iterations": sm["iterations"], "termination": sm["termination"],
            "final_cost": sm["final_cost"], "rot_err_rad": rot_angle(T[:3, :3], G[:3, :3]),
            "trans_err_m": float(np.linalg.norm(T[:3, 3] - G[:3, 3])), "launches": pnp.launch_count,
            "note": "planted extrinsic = reference config/pointgrey.bin, 2-D corners + N(0, 0.5 px); errors are the "
                    "method's accuracy on this synthetic scene, not a parity figure"}


def cpu_baseline(pts, clicks, sample_frames: int, threads: int):
    from oracle import oracle as O
    n = min(sample_frames, pts.shape[0])
    t = time.time()
    st, _ = O.extract_batch(pts[:n], clicks[:n], flags=O.FLAG_AUTO, threads=threads)
    dt = time.time() - t
    return n / dt, n, int((st == 0).sum())


def workload_text(frames: int) -> str:
    """config.workload: the same string on both arms (the driver compares them)."""
    return (f"{frames} synthetic VLP-16 frames (16x1800 XYZI, 28800 pts, packed 16 B) per GPU per step, "
            "7x5 board @0.15 m, colour phase auto (2 hypotheses x 2 solver phases per frame)")


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the real thing cannot be built here) on all
    host threads, same metric / config; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    sample = args.ref_sample_frames
    pts, clicks, _ = make_workload(sample, 0)
    for _ in range(args.warmup):
        O.extract_batch(pts[: max(threads, 16)], clicks[: max(threads, 16)], flags=O.FLAG_AUTO, threads=threads)
    t = time.time()
    for _ in range(args.steps):
        O.extract_batch(pts, clicks, flags=O.FLAG_AUTO, threads=threads)
    dt = time.time() - t
    fps = args.steps * sample / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.frames), "frames_per_gpu": args.frames,
                   "frames_total": args.gpus * args.frames, "seed": 20181203,
                   "reference_step": f"{sample} frames of that workload per step (bounded sample)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} frames x {args.steps} steps, OpenMP over frames"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--cpu-sample-frames", type=int, default=256)
    ap.add_argument("--ref-sample-frames", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU exchange: auto = fused into K5 (multicast if available, else peer stores), nccl = all_gather")
    ap.add_argument("--watchdog-s", type=int, default=900, help="dump all stacks and exit if the run exceeds this")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from lidar_camera_calibration_b200 import api, dist as ldist, synth
    rank, world, local = ldist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F, N, K, W = args.frames, synth.VLP16.n_points, args.steps, max(args.warmup, 3)

    pts, clicks, gt, gen_s = make_workload(F, rank * F, want_truth=True)
    ex = api.CornerExtractor(F, N, device=local)
    ncorn = ex.n_corners
    # device-resident inputs (value) and pinned host inputs (e2e)
    d_pts, d_clicks = torch.from_numpy(pts).to(dev), torch.from_numpy(clicks).to(dev)
    d_corners = torch.empty((F, ncorn, 3), dtype=torch.float32, device=dev)
    d_status = torch.empty((F,), dtype=torch.int32, device=dev)
    g_corners = torch.empty((world * F, ncorn, 3), dtype=torch.float32, device=dev) if world > 1 else None
    g_status = torch.empty((world * F,), dtype=torch.int32, device=dev) if world > 1 else None
    h_pts, h_clicks = torch.from_numpy(pts).pin_memory(), torch.from_numpy(clicks).pin_memory()
    h_corners = torch.empty((F, ncorn, 3), dtype=torch.float32).pin_memory()
    h_status = torch.empty((F,), dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    # multi-GPU exchange: K5 stores the corners straight into every rank's gather buffer over NVLink (one barrier per
    # step, no data-path collective); NCCL all-gather when symmetric memory is not available (--gather nccl forces it)
    fused = None
    if world > 1 and args.gather != "nccl":
        try:
            fused = ldist.FusedGather(ex, F, ncorn, rank, world, dev, use_multicast=(args.gather != "peer"))
            g_corners, g_status = fused.corners, fused.status
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] fused gather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
            fused = None
    if world > 1:
        flag = torch.tensor([1 if fused is not None else 0], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0 and fused is not None:
            fused.close()
            fused = None
            g_corners = torch.empty((world * F, ncorn, 3), dtype=torch.float32, device=dev)
            g_status = torch.empty((world * F,), dtype=torch.int32, device=dev)

    def step_device():
        ex.extract_torch(d_pts, d_clicks, d_corners, d_status, flags=api.FLAG_AUTO, stream=stream)
        if fused is not None:
            fused.barrier()
        elif world > 1:
            ldist.allgather_corners(d_corners, d_status, world, g_corners, g_status)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ----
    for _ in range(W):
        step_device()
    barrier()
    ex.set_profiling(True)
    stage_ms = {s: 0.0 for s in api.STAGES}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ex.launch_count
    e0.record(stream)
    for _ in range(K):
        step_device()
        for s, ms in ex.stage_times().items():  # per-kernel CUDA events on the same stream (syncs on the step's end)
            stage_ms[s] += ms
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    ex.set_profiling(False)
    launches = ex.launch_count - launches0
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    n_ok = int((d_status == 0).sum().item())

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H every step) ----
    def step_e2e():
        ex.extract_raw(h_pts.data_ptr(), F, N, 4, h_clicks.data_ptr(), h_corners.data_ptr(), h_status.data_ptr(),
                       api.MEM_HOST, api.FLAG_AUTO, stream.cuda_stream)

    Ke = max(3, min(K, 10))
    for _ in range(2):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(Ke):
        step_e2e()
        if fused is not None:
            fused.barrier()
        elif world > 1:
            ldist.allgather_corners(d_corners, d_status, world, g_corners, g_status)
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_match = bool(np.array_equal(h_corners.numpy(), d_corners.cpu().numpy(), equal_nan=True))

    # ---- configs[4]: the PnP that follows the all-gather (not part of the frames/s figure) ----
    step_device()
    torch.cuda.synchronize(dev)
    gather_ok = None
    if world > 1:  # whichever exchange ran, every rank must hold what a plain NCCL all-gather delivers, bit for bit
        ref_c, ref_s = ldist.allgather_corners(d_corners, d_status, world)
        same = torch.equal(ref_s, g_status) and torch.equal(torch.nan_to_num(ref_c, nan=-1.0), torch.nan_to_num(g_corners, nan=-1.0))
        t = torch.tensor([1 if same else 0], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        gather_ok = bool(t.item())
    pnp = pnp_leg(api, ldist, synth, torch, dev, rank, world, rank * F, gt,
                  g_corners if world > 1 else d_corners, g_status if world > 1 else d_status)

    # ---- projecting the batch's points through the extrinsic (spaceToPlane + colouring) ----
    proj = None
    if rank == 0:
        pj = api.PnpSolver(local)
        flat = d_pts.view(-1, 4)
        pix = torch.empty((flat.shape[0], 2), dtype=torch.int32, device=dev)
        rgbv = torch.empty((flat.shape[0], 4), dtype=torch.uint8, device=dev)
        for _ in range(3):
            pj.project_points(flat, synth.PLANTED_EXTRINSIC, synth.PLANTED_CAMERA, 1920, 1200, pix=pix, rgbv=rgbv)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(10):
            pj.project_points(flat, synth.PLANTED_EXTRINSIC, synth.PLANTED_CAMERA, 1920, 1200, pix=pix, rgbv=rgbv)
        p1.record(stream)
        torch.cuda.synchronize(dev)
        pms = p0.elapsed_time(p1) / 10
        pbytes = flat.shape[0] * (16 + 8 + 4)
        proj = {"points": int(flat.shape[0]), "ms": pms, "points_per_s": flat.shape[0] / (pms * 1e-3),
                "algorithmic_bytes": pbytes, "achieved_gbs": pbytes / (pms * 1e-3) / 1e9,
                "valid_points": int(rgbv[:, 3].sum().item()),
                "note": "ilcc2_project_points over the step's 1024 x 28800 points: 16 B in, 12 B out per point"}

    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    value = world * F * K / (elapsed_ms * 1e-3)
    e2e_value = world * F * Ke / (e2e_ms * 1e-3)
    peak, peak_kind = hbm_peak()
    dom = max(stage_ms, key=stage_ms.get)
    per_launch_ms = {s: stage_ms[s] / K for s in stage_ms}
    bytes_launch = BYTES_PER_FRAME * F

    def roof(ms):
        a = bytes_launch / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak}

    traffic = ncu_traffic() if F == 1024 else {}
    roofline = dict(roof(per_launch_ms[dom]), traffic=traffic.get(dom), kernel=dom, peak_source=f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                    algorithmic_bytes_per_launch=bytes_launch, kernel_ms_per_launch=per_launch_ms[dom],
                    share_of_step=stage_ms[dom] / sum(stage_ms.values()))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(F), "frames_per_gpu": F, "frames_total": world * F, "seed": synth.BASE_SEED,
                   "l2": f"inputs {F * N * 16 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                   "collective": ("none" if world == 1 else
                                  f"fused into K5 ({fused.mode} into symmetric memory) + one signal-pad barrier per step"
                                  if fused is not None else "nccl all_gather of corners+status per step"),
                   "frames_ok": n_ok, "gather_equals_nccl_all_gather": gather_ok},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * N * 16 + F * 12),
                "d2h_bytes_per_step": int(F * ncorn * 12 + F * 4), "steps": Ke, "ms_per_step": e2e_ms / Ke,
                "api": "ilcc2_extract_corners(mem_kind=ILCC2_MEM_HOST), pinned host buffers",
                "matches_device_path": e2e_match},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_path": dict(roof(elapsed_ms / K), note="whole step, all five kernels"),
        "roofline_k1": dict(roof(per_launch_ms["roi_filter"]), kernel="roi_filter", traffic=traffic.get("roi_filter"),
                            note="the only kernel that reads all input bytes"),
        "kernel_ms_per_step": per_launch_ms,
        "pnp": pnp,
        "projection": dict(proj, frac_of_hbm_peak=proj["achieved_gbs"] / peak) if proj else None,
        "workload_gen_s": gen_s,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = host_threads()
        v, n, ok = cpu_baseline(pts, clicks, args.cpu_sample_frames, threads)
        v1, n1, _ = cpu_baseline(pts, clicks, max(8, args.cpu_sample_frames // 16), 1)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"first {n} frames of the step's batch, OpenMP over frames ({ok} ok)",
                               "single_thread_value": v1, "single_thread_sample": f"first {n1} frames"}
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
