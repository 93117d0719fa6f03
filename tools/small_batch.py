"""Latency probe: times the stages for a small batch (one frame per SM) and the full batch."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
for F in (int(a) for a in sys.argv[1:]):
    pts, clicks = synth.make_batch(F)
    ex = api.CornerExtractor(F, synth.VLP16.n_points)
    dp, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(clicks).cuda()
    for _ in range(3):
        ex.extract_torch(dp, dc)
    torch.cuda.synchronize()
    ex.set_profiling(True)
    acc = {}
    for _ in range(5):
        ex.extract_torch(dp, dc)
        for k, v in ex.stage_times().items():
            acc[k] = acc.get(k, 0) + v / 5
    print(F, {k: round(v, 3) for k, v in acc.items()}, flush=True)
