#!/usr/bin/env python
"""bench.py -- LiDAR frames/s of the chessboard-corner hot path on synthetic 16x1800 XYZI clouds.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU path on the box's host cores)

A "step" is one pass of the whole hot path (K1 roi_filter, K2 cluster, K3+K4 plane_pca, K5 board_fit, + the exchange of
the corners between ranks when N>1: fused into K5 + one barrier, or an NCCL all-gather) over one batch of --frames
synthetic VLP-16 frames per GPU (BASELINE.json configs[2]; 8 ranks x 1024 frames = configs[4]).  ONE JSON line (rank 0):
  value            frames/s with the batch already resident in HBM (CUDA events around exactly K steps, max over ranks);
                   consecutive steps are taken in turn by --lanes handles on streams of their own, so step k+1's kernels
                   fill the tail of step k's K5; one_batch_in_flight is the same loop on one handle / one stream
  e2e              same metric through the C ABI with HOST buffers: pinned H2D of the batch + D2H of corners/status inside
                   the timed region; e2e_legs repeats it for pageable memory and for 32-byte PointCloud2 records
  roofline         dominant kernel of the step vs the measured HBM peak (algorithmic bytes = SURVEY.md 8(d));
                   roofline_compute: the issue-slot / fp64-pipe ceilings of the two latency-bound kernels (from the
                   committed ncu capture)
  parity           the step's batch compared frame for frame with the CPU oracle (outside the timed region)
  hdl64, single_frame_latency_ms, accuracy_vs_truth_m   BASELINE.json configs[3], [0]/[1]
  cpu_baseline     the CPU oracle (a restatement; the reference needs ROS/PCL/Ceres and cannot be built here) timed on
                   a bounded sample in the same run: -O3 -march=native build and the parity build
"""
from __future__ import annotations

import argparse
import glob
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
BYTES_PER_FRAME_HDL64 = 131072 * 16 + 12 + 63 * 12 + 4
HBM_FALLBACK_GBS = 6650.0                        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def newest_profile(pattern: str):
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    return paths[-1] if paths else None


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel, from the newest committed
    `ncu --set full` capture of this workload (1024 frames): profiles/ncu_traffic_*.json, written by
    tools/ncu_summary.py next to the text summaries. Returns ({stage: bytes}, file name)."""
    p = newest_profile("ncu_traffic_*.json")
    if not p:
        return {}, None
    try:
        return {k: v["dram_bytes"] for k, v in json.load(open(p)).items()}, os.path.basename(p)
    except Exception:
        return {}, None


def ncu_compute_ceilings():
    """roofline_compute: for the kernels that are NOT bandwidth-bound (K5 board_fit, K2 cluster) the ceilings they
    can be measured against: issue slots (warp instructions / (SMs x 4 schedulers x elapsed cycles)) and the fp64 pipe.
    Read from the newest committed ncu capture (profiles/ncu_compute_*.json, written by tools/ncu_summary.py)."""
    p = newest_profile("ncu_compute_*.json")
    if not p:
        return None
    try:
        d = json.load(open(p))
        d["source"] = "profiles/" + os.path.basename(p)
        return d
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.marks = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        self.marks.append(time.time())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.1)
        self.proc.terminate()
        t0, t1 = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, float("inf"))
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if not (t0 - 0.03 <= ts <= t1 + 0.03):
                continue
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
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "samples taken between the first and the last timed step (device-resident + end-to-end loops)"}


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


def cpu_rate(pts, clicks, n: int, threads: int, fast: bool):
    from oracle import oracle as O
    n = min(n, pts.shape[0])
    O.extract_batch(pts[:max(2, min(threads, n))], clicks[:max(2, min(threads, n))], flags=O.FLAG_AUTO, threads=threads, fast=fast)
    t = time.time()
    st, _ = O.extract_batch(pts[:n], clicks[:n], flags=O.FLAG_AUTO, threads=threads, fast=fast)
    return n / (time.time() - t), n, int((st == 0).sum())


def cpu_baseline_block(pts, clicks, sample_frames: int):
    """The CPU path on this box's host cores, same run: all cores (OpenMP over frames) and one thread (what the
    single-threaded reference does), for the -O3 -march=native build and for the parity build (-O2 -ffp-contract=off)."""
    threads = host_threads()
    v, n, ok = cpu_rate(pts, clicks, sample_frames, threads, True)
    v1, n1, _ = cpu_rate(pts, clicks, max(8, sample_frames // 32), 1, True)
    vp, _, _ = cpu_rate(pts, clicks, sample_frames, threads, False)
    return {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} frames of the step's batch, OpenMP over frames ({ok} ok), -O3 -march=native build",
            "single_thread_value": v1, "single_thread_sample": f"first {n1} frames",
            "parity_build_value": vp, "parity_build": "-O2 -ffp-contract=off (the build the parity tests use)"}


def workload_text(frames: int) -> str:
    """config.workload: the same string on both arms (the driver compares them)."""
    return (f"{frames} synthetic VLP-16 frames (16x1800 XYZI, 28800 pts, packed 16 B) per GPU per step, "
            "7x5 board @0.15 m, colour phase auto (2 hypotheses x 2 solver phases per frame)")


def base_config(args, world):
    return {"workload": workload_text(args.frames), "frames_per_gpu": args.frames, "frames_total": world * args.frames,
            "seed": 20181203}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the real thing cannot be built here) on all
    host threads, same metric / config: every step is the whole batch of --frames frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    F = args.frames
    pts, clicks, _ = make_workload(F, 0)
    for _ in range(max(1, min(args.warmup, 3))):
        O.extract_batch(pts[:max(threads * 4, 64)], clicks[:max(threads * 4, 64)], flags=O.FLAG_AUTO, threads=threads, fast=True)
    steps = max(1, min(args.steps, args.ref_max_steps))
    t = time.time()
    ok = 0
    for _ in range(steps):
        st, _ = O.extract_batch(pts, clicks, flags=O.FLAG_AUTO, threads=threads, fast=True)
        ok = int((st == 0).sum())
    dt = time.time() - t
    fps = steps * F / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": base_config(args, args.gpus),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{F} frames x {steps} steps (the whole batch per step; steps capped at --ref-max-steps "
                                   f"so the arm ends within minutes), OpenMP over frames, -O3 -march=native build, {ok} frames ok"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def hdl64_leg(api, synth, torch, dev, peak, frames=256, lanes=3):
    """BASELINE.json configs[3]: 64-ring x 2048-point clouds, 9x7 board (32 generated frames repeated to `frames`): one
    batch in flight (per-kernel times), and consecutive batches on `lanes` handles like the headline figure."""
    base = min(frames, 32)
    pts, clicks = synth.make_batch(base, scene=synth.SCENE_HDL64)
    reps = (frames + base - 1) // base
    pts, clicks = np.tile(pts, (reps, 1, 1))[:frames], np.tile(clicks, (reps, 1))[:frames]
    pipe = api.PipelinedExtractor(lanes, frames, synth.HDL64.n_points, device=dev.index,
                                  params=api.default_params(cells_x=8, cells_y=10), roi_capacity=32768)
    ex = pipe.extractors[0]
    ex.set_lanes(1)
    dp, dc = torch.from_numpy(pts).to(dev), torch.from_numpy(clicks).to(dev)
    outs_c = [torch.empty((frames, ex.n_corners, 3), dtype=torch.float32, device=dev) for _ in range(lanes)]
    outs_s = [torch.empty((frames,), dtype=torch.int32, device=dev) for _ in range(lanes)]
    for _ in range(2):
        st, _ = ex.extract_torch(dp, dc, outs_c[0], outs_s[0], flags=api.FLAG_AUTO)
    torch.cuda.synchronize(dev)
    R = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R):
        ex.extract_torch(dp, dc, outs_c[0], outs_s[0], flags=api.FLAG_AUTO)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / R
    ex.set_profiling(True)
    acc = {}
    for _ in range(3):
        ex.extract_torch(dp, dc, outs_c[0], outs_s[0], flags=api.FLAG_AUTO)
        for k, v in ex.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / 3
    ex.set_profiling(False)
    ref_c, ref_s = outs_c[0].clone(), outs_s[0].clone()
    ex.set_lanes(lanes)
    cur = torch.cuda.current_stream(dev)
    for _ in range(lanes):
        pipe.submit(dp, dc, outs_c, outs_s)
    pipe.join(cur)
    torch.cuda.synchronize(dev)
    Rl = 4 * lanes
    e0.record(cur)
    for _ in range(Rl):
        pipe.submit(dp, dc, outs_c, outs_s)
    pipe.join(cur)
    e1.record(cur)
    torch.cuda.synchronize(dev)
    ms_lanes = e0.elapsed_time(e1) / Rl
    same = all(torch.equal(s_, ref_s) and float((torch.nan_to_num(c_, nan=0.0) - torch.nan_to_num(ref_c, nan=0.0)).abs().max()) <= 1e-6
               for c_, s_ in zip(outs_c, outs_s))
    k1_gbs = BYTES_PER_FRAME_HDL64 * frames / (acc["roi_filter"] * 1e-3) / 1e9
    out = {"frames": frames, "points_per_frame": synth.HDL64.n_points, "board": "9x7 @0.15 m",
           "ms_per_step": ms_lanes, "frames_per_s": frames / (ms_lanes * 1e-3), "lanes": lanes,
           "one_batch_in_flight": {"ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3)},
           "lanes_match_one_batch": bool(same),
           "kernel_ms_per_step": acc, "frames_ok": int((st == 0).sum().item()),
           "k1_achieved_gbs": k1_gbs, "k1_frac_of_hbm_peak": k1_gbs / peak,
           "path_achieved_gbs": BYTES_PER_FRAME_HDL64 * frames / (ms_lanes * 1e-3) / 1e9,
           "note": "32 generated frames repeated; ROI 12-16 k points per frame -> K2 and K5 run on their large layouts; "
                   "kernel_ms_per_step: one batch in flight"}
    pipe.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--cpu-sample-frames", type=int, default=1024)
    ap.add_argument("--ref-max-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the hdl64 / single-frame / pnp / projection legs")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="multi-GPU exchange: auto = fused into K5 (multicast if available, else peer stores), nccl = all_gather")
    ap.add_argument("--lanes", type=int, default=3,
                    help="handles (each on its own CUDA stream) that take consecutive steps in turn; 1 = one batch in flight")
    ap.add_argument("--gather-parts", type=int, default=0, choices=[0, 2, 3],
                    help="parts of every lane's symmetric gather buffer: 3 = the signal-pad barrier runs on a side stream, one "
                         "step behind the lane (dist.FusedGather); 2 = inline on the lane's stream; 0 = 3 from four ranks on "
                         "(measured 1.78 -> 1.75 ms per step at N=4, 1.83 -> 1.78 at N=8; at N=2 inline is faster: 1.73 vs 1.77)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin ranks to the NUMA node of their GPU")
    ap.add_argument("--watchdog-s", type=int, default=900, help="dump all stacks and exit if the run exceeds this")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog_s, exit=True)
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries the ONE JSON line: whatever native libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import torch
    from lidar_camera_calibration_b200 import api, dist as ldist, synth
    rank, world, local = ldist.init_from_env()
    if args.gather_parts == 0:
        args.gather_parts = 3 if world >= 4 else 2
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one rank per GPU: keep this rank's threads and page-locked buffers on the NUMA node of its GPU (H2D copies of the
    # end-to-end legs then stay off the inter-socket link); a single rank keeps all cores (cpu_baseline uses them)
    numa = ldist.bind_to_gpu_numa(local) if world > 1 and not args.no_numa_bind else None
    F, N, K, W = args.frames, synth.VLP16.n_points, args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up, so that it is up when the timed region starts

    pts, clicks, gt, gen_s = make_workload(F, rank * F, want_truth=True)
    # `lanes` handles on streams of their own take the steps in turn (api.PipelinedExtractor): every step is the same
    # launches on the same batch, but step k+1's kernels fill the SMs that the tail of step k's K5 leaves idle
    L = max(1, args.lanes)
    pipe = api.PipelinedExtractor(L, F, N, device=local, roi_capacity=8192)
    ex = pipe.extractors[0]
    ncorn = ex.n_corners
    # device-resident inputs (value) and pinned host inputs (e2e)
    d_pts, d_clicks = torch.from_numpy(pts).to(dev), torch.from_numpy(clicks).to(dev)
    lane_corners = [torch.empty((F, ncorn, 3), dtype=torch.float32, device=dev) for _ in range(L)]
    lane_status = [torch.empty((F,), dtype=torch.int32, device=dev) for _ in range(L)]
    d_corners, d_status = lane_corners[0], lane_status[0]
    h_pts, h_clicks = torch.from_numpy(pts).pin_memory(), torch.from_numpy(clicks).pin_memory()
    h_corners = torch.empty((F, ncorn, 3), dtype=torch.float32).pin_memory()
    h_status = torch.empty((F,), dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    # multi-GPU exchange: K5 stores the corners straight into every rank's gather buffer over NVLink (one barrier per
    # step, no data-path collective); NCCL all-gather when symmetric memory is not available (--gather nccl forces it)
    fused_l = None   # one exchange object per lane: its own double-buffered symmetric buffer and signal pads
    if world > 1 and args.gather != "nccl":
        try:
            fused_l = [ldist.FusedGather(pipe.extractors[j], F, ncorn, rank, world, dev, use_multicast=(args.gather != "peer"),
                                         halves=args.gather_parts) for j in range(L)]
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] fused gather unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
            fused_l = None
    if world > 1:
        flag = torch.tensor([1 if fused_l is not None else 0], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if int(flag.item()) == 0 and fused_l is not None:
            for fg in fused_l:
                fg.close()
            fused_l = None
    fused = fused_l[0] if fused_l is not None else None
    nccl_gather = world > 1 and fused is None
    lane_gc = [torch.empty((world * F, ncorn, 3), dtype=torch.float32, device=dev) for _ in range(L)] if nccl_gather else None
    lane_gs = [torch.empty((world * F,), dtype=torch.int32, device=dev) for _ in range(L)] if nccl_gather else None
    g_corners, g_status = (lane_gc[0], lane_gs[0]) if nccl_gather else (None, None)

    def gathered():  # for a consumer on the current stream
        if fused is not None:
            fused.wait_ready()
            return fused.corners, fused.status
        return g_corners, g_status

    def exchanges_done(s):  # `s` waits for the last barrier of every lane (they may run on side streams)
        if fused_l is not None:
            for fg in fused_l:
                fg.wait_ready(s)

    def exchange(j=0):  # on the current stream
        if fused_l is not None:
            fused_l[j].barrier()
        elif world > 1:
            ldist.allgather_corners(lane_corners[j], lane_status[j], world, lane_gc[j], lane_gs[j])

    def step_device():
        pipe.submit(d_pts, d_clicks, lane_corners, lane_status, flags=api.FLAG_AUTO, after=exchange if world > 1 else None)

    def step_one_lane():
        ex.extract_torch(d_pts, d_clicks, d_corners, d_status, flags=api.FLAG_AUTO, stream=stream)
        exchange(0)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput: exactly K steps, nothing else inside the timed region ----
    def timed(step, steps):
        for _ in range(W):
            step()
        pipe.join(stream)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = pipe.launch_count
        sampler.mark()
        t0.record(stream)     # the lanes' streams wait for `stream` at every submit, i.e. they start behind t0
        for _ in range(steps):
            step()
        pipe.join(stream)     # `stream` waits for the last step of every lane ...
        exchanges_done(stream)  # ... and for its exchange
        t1.record(stream)
        barrier()
        sampler.mark()
        return max_over_ranks(t0.elapsed_time(t1)), pipe.launch_count - n0

    elapsed_ms, launches = timed(step_device, K)
    n_ok = int((d_status == 0).sum().item())
    lanes_agree = all(torch.equal(lane_status[j], d_status) and
                      torch.equal(torch.nan_to_num(lane_corners[j], nan=-1.0), torch.nan_to_num(d_corners, nan=-1.0)) for j in range(L))
    # the same K steps with ONE batch in flight (one handle, one stream): what the overlap of consecutive steps is worth
    ex.set_lanes(1)
    one_lane_ms = timed(step_one_lane, K)[0] if L > 1 else elapsed_ms

    # ---- per-kernel times: a separate pass with the library's CUDA events around every stage (same stream) ----
    def stage_pass():
        ex.set_profiling(True)
        acc = {s: 0.0 for s in api.STAGES}
        for _ in range(Kp):
            ex.extract_torch(d_pts, d_clicks, d_corners, d_status, flags=api.FLAG_AUTO, stream=stream)
            for s, ms in ex.stage_times().items():
                acc[s] += ms / Kp
        ex.set_profiling(False)
        return acc

    Kp = max(5, min(K, 20))
    stage_ms = stage_pass()            # one batch in flight: the launch shapes of a single handle (K5: 2 warps per solve)
    ex.set_lanes(L)
    stage_ms_lane = stage_pass() if L > 1 else stage_ms   # the shapes the lanes run (K5: 1 warp per solve), timed alone
    exchange()  # keeps the ranks' double buffers in step

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H every step) ----
    h_corners_l = [h_corners] + [torch.empty((F, ncorn, 3), dtype=torch.float32).pin_memory() for _ in range(L - 1)]
    h_status_l = [h_status] + [torch.empty((F,), dtype=torch.int32).pin_memory() for _ in range(L - 1)]

    def e2e_run(points_t, width, steps, lanes=L):
        """The lanes take the steps in turn, each with a host thread in the blocking C-ABI call (PipelinedExtractor.submit_host);
        a step's exchange is enqueued on its lane's stream once its call has returned."""
        pending = [False] * L

        def finish(j):  # step of lane j: call returned -> its exchange (this thread, lane j's stream)
            pipe.wait_lane(j)
            if pending[j] and world > 1:
                with torch.cuda.stream(pipe.streams[j]):
                    if fused_l is not None:
                        fused_l[j].barrier()
                    else:  # the call left this step's corners in the pinned host buffers: upload them, then gather
                        lane_corners[j].copy_(h_corners_l[j], non_blocking=True)
                        lane_status[j].copy_(h_status_l[j], non_blocking=True)
                        ldist.allgather_corners(lane_corners[j], lane_status[j], world, lane_gc[j], lane_gs[j])
            pending[j] = False

        def run(n):
            for _ in range(n):
                j = pipe.next_lane
                finish(j)
                pipe.submit_host(points_t.data_ptr(), F, N, width, h_clicks.data_ptr(), h_corners_l[j].data_ptr(),
                                 h_status_l[j].data_ptr(), api.FLAG_AUTO)
                pending[j] = True
            for j in range(L):
                finish(j)
            pipe.join(stream)
            for j in range(L):
                stream.wait_stream(pipe.streams[j])
            exchanges_done(stream)

        pipe.active = lanes
        run(2 * lanes)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        run(steps)
        f1.record(stream)
        barrier()
        pipe.active = L
        return max_over_ranks(f0.elapsed_time(f1)) / steps

    Ke = max(3, min(K, 10))
    sampler.mark()
    e2e_ms = e2e_run(h_pts, 4, Ke)
    sampler.mark()
    e2e_one_ms = e2e_run(h_pts, 4, Ke, lanes=1) if L > 1 else e2e_ms

    def h2d_only(steps):
        """The floor of the end-to-end figure on THIS box: nothing but the step's H2D copy (page-locked -> device), all
        ranks copying at the same time like in the end-to-end loop."""
        scratch = torch.empty_like(d_pts)
        scratch.copy_(h_pts, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(steps):
            scratch.copy_(h_pts, non_blocking=True)
        c1.record(stream)
        barrier()
        return max_over_ranks(c0.elapsed_time(c1)) / steps

    h2d_ms = h2d_only(Ke)
    e2e_match = bool(np.array_equal(h_corners.numpy(), d_corners.cpu().numpy(), equal_nan=True))
    clocks = sampler.stop() if rank == 0 else None
    legs = {"packed16_pinned": {"ms_per_step": e2e_ms, "frames_per_s": world * F / (e2e_ms * 1e-3),
                                "h2d_bytes_per_step": int(F * N * 16 + F * 12)}}
    if not args.no_extras:
        # the reference's real input form: the `data` vector of a sensor_msgs::PointCloud2 -- pageable memory, 32-byte
        # records x@0 y@4 z@8 intensity@16 (get_lidar_corners.cpp:136-164) -- and packed records in pageable memory
        rec32 = np.zeros((F, N, 8), np.float32)
        rec32[:, :, 0:3] = pts[:, :, 0:3]
        rec32[:, :, 4] = pts[:, :, 3]
        for name, arr, width in (("packed16_pageable", torch.from_numpy(pts), 4),
                                 ("stride32_pinned", torch.from_numpy(rec32).pin_memory(), 8),
                                 ("stride32_pageable", torch.from_numpy(rec32), 8)):
            ms = e2e_run(arr, width, max(3, Ke // 2))
            legs[name] = {"ms_per_step": ms, "frames_per_s": world * F / (ms * 1e-3),
                          "h2d_bytes_per_step": int(F * N * width * 4 + F * 12),
                          "matches_device_path": bool(np.array_equal(h_corners.numpy(), d_corners.cpu().numpy(), equal_nan=True))}
        legs["pageable_over_pinned"] = {"packed16": legs["packed16_pageable"]["ms_per_step"] / legs["packed16_pinned"]["ms_per_step"],
                                        "stride32": legs["stride32_pageable"]["ms_per_step"] / legs["stride32_pinned"]["ms_per_step"]}
        del rec32

    # ---- the exchange delivers what a plain NCCL all-gather delivers, also when consecutive steps differ ----
    gather_ok = None
    if world > 1:
        same = True
        for shift in (0, 1, 2):  # three consecutive steps on different inputs (frames rolled by `shift`)
            p, c = (d_pts, d_clicks) if shift == 0 else (torch.roll(d_pts, shift, 0), torch.roll(d_clicks, shift, 0))
            ex.extract_torch(p, c, d_corners, d_status, flags=api.FLAG_AUTO, stream=stream)
            exchange()
            gc, gs = gathered()
            ref_c, ref_s = ldist.allgather_corners(d_corners, d_status, world)
            same = same and torch.equal(ref_s, gs) and torch.equal(torch.nan_to_num(ref_c, nan=-1.0), torch.nan_to_num(gc, nan=-1.0))
            if fused is not None:
                fused.release()
        t = torch.tensor([1 if same else 0], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        gather_ok = bool(t.item())
    # ---- configs[4]: the PnP that follows the all-gather (not part of the frames/s figure) ----
    step_device()
    torch.cuda.synchronize(dev)
    pnp = None
    if not args.no_extras:
        gc, gs = gathered() if world > 1 else (d_corners, d_status)
        pnp = pnp_leg(api, ldist, synth, torch, dev, rank, world, rank * F, gt, gc, gs)
    if rank != 0:
        torch.distributed.destroy_process_group()
        return

    peak, peak_kind = hbm_peak()
    # ---- projecting the batch's points through the extrinsic (spaceToPlane + colouring) ----
    proj = hdl = None
    single_ms = None
    if not args.no_extras:
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
                "frac_of_hbm_peak": pbytes / (pms * 1e-3) / 1e9 / peak, "valid_points": int(rgbv[:, 3].sum().item()),
                "note": "ilcc2_project_points over the step's 1024 x 28800 points: 16 B in, 12 B out per point"}
        del pix, rgbv
        # configs[0]/[1]: one frame at a time (device-resident, 8 warps per solve): latency of a single call
        ex1 = api.CornerExtractor(1, N, device=local, roi_capacity=8192)
        c1, s1 = torch.empty((1, ncorn, 3), dtype=torch.float32, device=dev), torch.empty((1,), dtype=torch.int32, device=dev)
        for _ in range(3):
            ex1.extract_torch(d_pts[:1], d_clicks[:1], c1, s1, flags=api.FLAG_AUTO, stream=stream)
        torch.cuda.synchronize(dev)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        for i in range(20):
            ex1.extract_torch(d_pts[i:i + 1], d_clicks[i:i + 1], c1, s1, flags=api.FLAG_AUTO, stream=stream)
        q1.record(stream)
        torch.cuda.synchronize(dev)
        single_ms = q0.elapsed_time(q1) / 20
        ex1.close()
        try:
            hdl = hdl64_leg(api, synth, torch, dev, peak, lanes=L)
        except Exception as e:  # noqa: BLE001
            hdl = {"error": f"{type(e).__name__}: {e}"}

    # ---- accuracy against the planted ground truth (configs[1]) and parity against the oracle (whole batch) ----
    st_h, c_h, diags = ex.extract(pts, clicks, flags=api.FLAG_AUTO, want_diag=True)   # host path, diag records
    okf = np.flatnonzero(st_h == 0)
    acc = np.array([synth.corner_set_distance(c_h[f], gt[f]) for f in okf]) if okf.size else np.zeros(1)
    accuracy = {"frames": int(okf.size), "median": float(np.median(acc)), "p95": float(np.percentile(acc, 95)),
                "frac_within_1cm": float((acc < 1e-2).mean()),
                "note": "max corner distance per frame to the planted corners (order-normalised): the METHOD's accuracy on "
                        "this scene (the reference's optimiser settles one square off on a minority of frames), not parity"}
    parity = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        parity = O.parity_report(pts, clicks, st_h, c_h, diags, threads=host_threads())
        parity["note"] = ("oracle (parity build) on ALL frames of the step's batch vs the C ABI's results; corners compared "
                          "index for index, tolerance = north_star's 1e-3 m")
        c_d = d_corners.cpu().numpy()
        parity["host_path_equals_device_path"] = bool(np.array_equal(c_h, c_d, equal_nan=True))
        both = np.isfinite(c_h).all(axis=(1, 2)) & np.isfinite(c_d).all(axis=(1, 2))
        parity["host_vs_device_max_diff_m"] = float(np.abs(c_h[both] - c_d[both]).max()) if both.any() else None

    value = world * F * K / (elapsed_ms * 1e-3)
    e2e_value = world * F / (e2e_ms * 1e-3)
    dom = max(stage_ms, key=stage_ms.get)
    bytes_launch = BYTES_PER_FRAME * F

    def roof(ms):
        a = bytes_launch / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak}

    traffic, traffic_src = ncu_traffic() if F == 1024 else ({}, None)
    tkey = {"plane_pca": "plane_pca"}.get(dom, dom)
    roofline = dict(roof(stage_ms[dom]), traffic=traffic.get(tkey), traffic_source=traffic_src, kernel=dom,
                    peak_source=f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                    algorithmic_bytes_per_launch=bytes_launch, kernel_ms_per_launch=stage_ms[dom],
                    share_of_step=stage_ms[dom] / sum(stage_ms.values()))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(base_config(args, world),
                       l2=f"inputs {F * N * 16 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       collective=("none" if world == 1 else
                                   f"fused into K5 ({fused.mode} into {args.gather_parts}-part symmetric memory per lane) + one signal-pad barrier per step"
                                   + (" on a side stream, one step behind the lane" if args.gather_parts == 3 else "")
                                   if fused is not None else "nccl all_gather of corners+status per step"),
                       frames_ok=n_ok, gather_equals_nccl_all_gather=gather_ok, numa_binding=numa,
                       pipeline=f"{L} handles x {L} CUDA streams take the K steps in turn (api.PipelinedExtractor); every step = the "
                                "same launches on the same batch, consecutive steps overlap on the device" if L > 1 else
                                "one handle, one stream: one batch in flight"),
        "one_batch_in_flight": {"value": world * F * K / (one_lane_ms * 1e-3), "ms_per_step": one_lane_ms / K, "unit": UNIT,
                                "note": "the same K steps on one handle and one stream (no overlap of consecutive steps); "
                                        "kernel_ms_per_step and the rooflines are measured this way",
                                "lanes_agree": lanes_agree},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(F * N * 16 + F * 12),
                "d2h_bytes_per_step": int(F * ncorn * 12 + F * 4), "steps": Ke, "ms_per_step": e2e_ms,
                "api": "ilcc2_extract_corners(mem_kind=ILCC2_MEM_HOST), pinned host buffers, packed 16-byte records; "
                       f"{L} handles, each with a host thread in the blocking call, take the steps in turn",
                "matches_device_path": e2e_match,
                "one_call_in_flight": {"ms_per_step": e2e_one_ms, "value": world * F / (e2e_one_ms * 1e-3)},
                "h2d_copy_only_ms": h2d_ms, "h2d_copy_only_gbs_per_gpu": F * N * 16 / (h2d_ms * 1e-3) / 1e9,
                "h2d_copy_only_note": "the step's 472 MB cudaMemcpyAsync from page-locked memory and nothing else, all ranks at "
                                      "the same time: the floor of e2e on this box (the ranks share the host's PCIe / memory path)"},
        "e2e_legs": legs,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_path": dict(roof(elapsed_ms / K), note="whole step, all kernels"),
        "roofline_k1": dict(roof(stage_ms["roi_filter"]), kernel="roi_filter", traffic=traffic.get("roi_filter"),
                            note="the only kernel that reads all input bytes"),
        "roofline_compute": ncu_compute_ceilings(),
        "kernel_ms_per_step": stage_ms,
        "kernel_ms_note": "CUDA events of the library around each stage, a separate pass of the same workload with ONE batch in "
                          "flight (the timed region of `value` holds nothing but the K steps)",
        "kernel_ms_per_step_lane_shapes": stage_ms_lane,
        "kernel_ms_lane_shapes_note": "the same pass with the launch shapes the lanes use (ilcc2_set_lanes: K5 with one warp per "
                                      "solve -- slower for a batch on its own, faster when consecutive batches overlap)",
        "parity": parity,
        "accuracy_vs_truth_m": accuracy,
        "single_frame_latency_ms": single_ms,
        "hdl64": hdl,
        "pnp": pnp,
        "projection": proj,
        "scratch_mb": round(ex_scratch_mb(F, 8192, ncorn), 1), "staging_ring_mb": round(staging_mb(F), 1), "handles": L,
        "workload_gen_s": gen_s,
    }
    if not args.no_cpu_baseline:  # rank 0, at every N (with N>1 it is bound to one NUMA node: `cores` says how many)
        out["cpu_baseline"] = cpu_baseline_block(pts, clicks, args.cpu_sample_frames)
    emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


def ex_scratch_mb(F, roi_cap, ncorn):
    """Device scratch of a handle sized with roi_capacity (ilcc2_create_sized), MB: ROI-sized arrays (41 B per entry),
    the pool of global-memory clustering tables, diag records (the staging ring for host batches: staging_mb)."""
    ts = 1
    while ts < 2 * roi_cap:
        ts <<= 1
    pool = F if roi_cap > 4096 else max(min(F, 8), F // 8)
    return (F * roi_cap * 41 + pool * (2 * ts + 3 * roi_cap) * 4 + F * (320 + ncorn * 12 + 64)) / 1e6


def staging_mb(F):
    """The handle's device staging ring for host batches: 4 slots of 64 frames of 32-byte records."""
    return 4 * min(F, 64) * 28800 * 32 / 1e6


if __name__ == "__main__":
    main()
