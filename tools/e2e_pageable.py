"""End-to-end ms per 1024-frame step through ilcc2_extract_corners(MEM_HOST) for pinned / pageable caller memory and
both record layouts; env knobs (ILCC2_STAGE_THREADS, ILCC2_DRIVER_STAGING, ILCC2_E2E_CHUNK) are A/B'd by the caller."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
F, N = 1024, synth.VLP16.n_points
cache = "/tmp/ab_inputs_1024.npz"
if os.path.exists(cache):
    z = np.load(cache); pts, clicks = z["pts"], z["clicks"]
else:
    pts, clicks = synth.make_batch(F); np.savez(cache, pts=pts, clicks=clicks)
rec32 = np.zeros((F, N, 8), np.float32); rec32[:, :, :3] = pts[:, :, :3]; rec32[:, :, 4] = pts[:, :, 3]
ex = api.CornerExtractor(F, N, roi_capacity=8192)
hc = torch.empty((F, 35, 3)).pin_memory(); hs = torch.empty((F,), dtype=torch.int32).pin_memory()
hk = torch.from_numpy(clicks).pin_memory()
st = torch.cuda.current_stream()
out = []
for name, t, w in (("p16_pin", torch.from_numpy(pts).pin_memory(), 4), ("p16_page", torch.from_numpy(pts), 4),
                   ("s32_pin", torch.from_numpy(rec32).pin_memory(), 8), ("s32_page", torch.from_numpy(rec32), 8)):
    def one():
        ex.extract_raw(t.data_ptr(), F, N, w, hk.data_ptr(), hc.data_ptr(), hs.data_ptr(), api.MEM_HOST, api.FLAG_AUTO, st.cuda_stream)
    for _ in range(2): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): one()
    e1.record(); torch.cuda.synchronize()
    out.append(f"{name}={e0.elapsed_time(e1)/5:.2f}")
print(" ".join(f"{k}={os.environ[k]}" for k in ("ILCC2_STAGE_THREADS", "ILCC2_DRIVER_STAGING", "ILCC2_E2E_CHUNK") if k in os.environ) or "default", "|", " ".join(out), flush=True)
