"""End-to-end (host buffers through the C ABI) with L handles driven by L host threads taking the batches in turn:
python tools/e2e_lanes.py [frames] [steps]"""
import sys, threading, time, queue
sys.path.insert(0, '.')
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 12
N = synth.VLP16.n_points
pts, clicks = synth.make_batch(F)
PAGEABLE = len(sys.argv) > 3 and sys.argv[3] == 'pageable'
h_pts = torch.from_numpy(pts) if PAGEABLE else torch.from_numpy(pts).pin_memory()
h_clicks = torch.from_numpy(clicks).pin_memory()
print('pageable' if PAGEABLE else 'pinned')
ref = None
for L in (1, 2, 3):
    exs = [api.CornerExtractor(F, N, roi_capacity=8192) for _ in range(L)]
    outs = [(torch.empty((F, exs[0].n_corners, 3), dtype=torch.float32).pin_memory(), torch.empty((F,), dtype=torch.int32).pin_memory()) for _ in range(L)]
    streams = [torch.cuda.Stream() for _ in range(L)]
    qs = [queue.Queue() for _ in range(L)]
    def worker(j):
        torch.cuda.set_device(0)
        while True:
            it = qs[j].get()
            if it is None:
                return
            exs[j].extract_raw(h_pts.data_ptr(), F, N, 4, h_clicks.data_ptr(), outs[j][0].data_ptr(), outs[j][1].data_ptr(),
                               api.MEM_HOST, api.FLAG_AUTO, streams[j].cuda_stream)
            qs[j].task_done()
    th = [threading.Thread(target=worker, args=(j,), daemon=True) for j in range(L)]
    for t in th: t.start()
    def run(steps):
        for i in range(steps):
            qs[i % L].put(1)
        for q in qs: q.join()
    run(2 * L)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(K)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / K
    if ref is None:
        ref = (outs[0][0].numpy().copy(), outs[0][1].numpy().copy())
    same = all(np.array_equal(o[0].numpy(), ref[0], equal_nan=True) and np.array_equal(o[1].numpy(), ref[1]) for o in outs)
    print(f'lanes {L}: {ms:.3f} ms/step  {F / ms * 1e3:,.0f} frames/s  identical: {same}', flush=True)
    for q in qs: q.put(None)
    for e in exs: e.close()
