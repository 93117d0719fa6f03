"""End-to-end (pinned host buffers -> corners on the host) time of the bench workload for several ILCC2_E2E_CHUNK values.
    python tools/e2e_ab.py 128 192 256 384 512"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "--child":
    import numpy as np, torch
    from lidar_camera_calibration_b200 import api, synth
    F = 1024
    cache = f"/tmp/ab_inputs_{F}.npz"
    if os.path.exists(cache):
        z = np.load(cache); pts, clicks = z["pts"], z["clicks"]
    else:
        pts, clicks = synth.make_batch(F); np.savez(cache, pts=pts, clicks=clicks)
    ex = api.CornerExtractor(F, synth.VLP16.n_points)
    hp, hc = torch.from_numpy(pts).pin_memory(), torch.from_numpy(clicks).pin_memory()
    ho = torch.empty((F, ex.n_corners, 3), dtype=torch.float32).pin_memory(); hs = torch.empty((F,), dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream()
    def step():
        ex.extract_raw(hp.data_ptr(), F, synth.VLP16.n_points, 4, hc.data_ptr(), ho.data_ptr(), hs.data_ptr(), api.MEM_HOST, api.FLAG_AUTO, st.cuda_stream)
    for _ in range(3): step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); R = 10
    for _ in range(R): step()
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {sys.argv[2]:>5s}: e2e {e0.elapsed_time(e1) / R:.3f} ms per 1024 frames, ok {int((hs == 0).sum())}", flush=True)
    sys.exit(0)
for c in sys.argv[1:]:
    subprocess.run([sys.executable, __file__, "--child", c], env=dict(os.environ, ILCC2_E2E_CHUNK=c))
