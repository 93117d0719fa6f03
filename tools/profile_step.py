"""Profiling driver: the bench workload (1024 synthetic frames, device-resident), `steps` passes of the hot path.
    ncu --set full --clock-control none --import-source on -k regex:board_fit -s 1 -c 1 -o rep python tools/profile_step.py 3"""
import sys
sys.path.insert(0, '.')
import torch
from lidar_camera_calibration_b200 import api, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
F = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pts, clicks = synth.make_batch(F)
ex = api.CornerExtractor(F, synth.VLP16.n_points, roi_capacity=8192)
if len(sys.argv) > 3:
    ex.set_lanes(int(sys.argv[3]))   # launch shapes of a lane of the throughput mode (K5: one warp per solve with 3 lanes)
dp, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(clicks).cuda()
for _ in range(steps):
    st, c = ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
torch.cuda.synchronize()
print('ok', int((st == 0).sum().item()), 'of', F, flush=True)
