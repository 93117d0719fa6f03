"""BASELINE.json configs[3]: 64-ring x 2048 HDL-64-like frames with a 9x7 board (global-memory code paths of K2 / K5).
    python tools/hdl64_bench.py [frames]   -> per-stage ms, frames/s, GB/s of algorithmic bytes"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = time.time()
pts, clicks = synth.make_batch(min(F, 32), scene=synth.SCENE_HDL64)
reps = (F + pts.shape[0] - 1) // pts.shape[0]
pts = np.ascontiguousarray(np.tile(pts, (reps, 1, 1))[:F]); clicks = np.ascontiguousarray(np.tile(clicks, (reps, 1))[:F])
print(f"generated {min(F, 32)} distinct frames (tiled to {F}) in {time.time() - t:.1f} s", flush=True)
ex = api.CornerExtractor(F, synth.HDL64.n_points, params=api.default_params(cells_x=8, cells_y=10), roi_capacity=32768)
dp, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(clicks).cuda()
for _ in range(3):
    st, c = ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
torch.cuda.synchronize()
ex.set_profiling(True)
acc, R = {}, 5
for _ in range(R):
    ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
    for k, v in ex.stage_times().items():
        acc[k] = acc.get(k, 0) + v / R
ex.set_profiling(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(R):
    ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
st, c, d = ex.extract(pts[:4], clicks[:4], flags=api.FLAG_AUTO, want_diag=True)
bpf = 131072 * 16 + 12 + 63 * 12 + 4
print(f"F={F} ok={int((st == 0).sum())}/4 n_roi={[x['n_roi'] for x in d]} n_opt={[x['n_opt'] for x in d]}")
print(f"step {ms:.3f} ms = {F / ms * 1e3:.0f} frames/s = {F * bpf / ms / 1e6:.1f} GB/s algorithmic; stages " +
      " ".join(f"{k}={v:.3f}" for k, v in acc.items()), flush=True)
