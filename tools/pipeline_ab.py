"""Throughput of consecutive device-resident batches when D handles on D streams take the batches in turn (so the tail of
one batch's K5 is filled by the next batch's kernels):  python tools/pipeline_ab.py [frames] [steps]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device('cuda', 0)
pts, clicks = synth.make_batch(F)
d_pts, d_clicks = torch.from_numpy(pts).to(dev), torch.from_numpy(clicks).to(dev)
ref = None
CASES = ((1, False), (2, False)) if os.environ.get('PIPE_SHORT') else ((1, False), (2, False), (2, True), (3, False), (4, False))
if os.environ.get('PIPE_CASES'):
    CASES = tuple((int(x), False) for x in os.environ['PIPE_CASES'].split(','))
for D, prio in CASES:
    exs = [api.CornerExtractor(F, synth.VLP16.n_points, roi_capacity=8192) for _ in range(D)]
    for e in exs:
        e.set_lanes(int(os.environ.get('PIPE_SET_LANES', D)))
    streams = [torch.cuda.Stream(dev, priority=(-1 if (prio and i % 2 == 0) else 0)) for i in range(D)]
    outs = [(torch.empty((F, exs[0].n_corners, 3), dtype=torch.float32, device=dev), torch.empty((F,), dtype=torch.int32, device=dev)) for _ in range(D)]
    def run(steps):
        for i in range(steps):
            j = i % D
            exs[j].extract_torch(d_pts, d_clicks, outs[j][0], outs[j][1], flags=api.FLAG_AUTO, stream=streams[j])
    run(2 * D)
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    for s in streams:
        s.wait_event(e0)
    run(K)
    ends = []
    for s in streams:
        e = torch.cuda.Event(enable_timing=True); e.record(s); ends.append(e)
    torch.cuda.synchronize(dev)
    ms = max(e0.elapsed_time(e) for e in ends) / K
    c = outs[0][0].cpu().numpy(); st = outs[0][1].cpu().numpy()
    if ref is None:
        ref = (c, st)
    same = all(np.array_equal(o[0].cpu().numpy(), ref[0], equal_nan=True) and np.array_equal(o[1].cpu().numpy(), ref[1]) for o in outs)
    print(f'handles {D} prio {prio}: {ms:.3f} ms/step  {F / ms * 1e3:,.0f} frames/s  outputs identical to depth 1: {same}', flush=True)
    for e in exs: e.close()
