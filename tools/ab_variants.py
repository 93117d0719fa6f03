"""A/B of build variants (build_variants/lib_*.so) on the bench workload: per-stage ms and agreement of the results
with the first variant.  usage: python tools/ab_variants.py [frames] name1 name2 ...   (child mode: --child name)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "--child":
    import numpy as np, torch
    from lidar_camera_calibration_b200 import api, synth
    name, F = sys.argv[2], int(sys.argv[3])
    cache = f"/tmp/ab_inputs_{F}.npz"
    if os.path.exists(cache):
        z = np.load(cache); pts, clicks = z["pts"], z["clicks"]
    else:
        pts, clicks = synth.make_batch(F)
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.savez(cache, pts=pts, clicks=clicks)
    ex = api.CornerExtractor(F, synth.VLP16.n_points)
    dp, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(clicks).cuda()
    for _ in range(3):
        ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
    torch.cuda.synchronize()
    ex.set_profiling(True)
    acc, R = {}, 8
    for _ in range(R):
        st, c = ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
        for k, v in ex.stage_times().items():
            acc[k] = acc.get(k, 0) + v / R
    ex.set_profiling(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R):
        ex.extract_torch(dp, dc, flags=api.FLAG_AUTO)
    e1.record(); torch.cuda.synchronize()
    st, c, d = ex.extract(pts, clicks, flags=api.FLAG_AUTO, want_diag=True)
    it = np.array([x["iters"] for x in d]).reshape(F, -1); th = np.array([x["theta_t"] for x in d]).reshape(F, -1)
    ref = f"/tmp/ab_ref_{F}.npz"
    msg = ""
    if os.path.exists(ref):
        z = np.load(ref)
        ok = (st == 0) & (z["st"] == 0)
        msg = (f" | vs ref: status_eq={np.array_equal(st, z['st'])} iters_eq={np.array_equal(it, z['it'])} "
               f"max|dcorner|={np.abs(c[ok] - z['c'][ok]).max():.2e} max|dtheta_t|={np.abs(th[ok] - z['th'][ok]).max():.2e}")
    else:
        np.savez(ref, st=st, c=c, it=it, th=th)
    print(f"{name:>28s} F={F} step={e0.elapsed_time(e1) / R:.3f} ms  " + " ".join(f"{k}={v:.3f}" for k, v in acc.items()) + msg, flush=True)
    sys.exit(0)
F = int(sys.argv[1])
for name in sys.argv[2:]:
    # name[@ENV=VAL[@ENV=VAL...]]; the name "tree" is the in-tree build
    lib, *assign = name.split("@")
    path = os.path.join(ROOT, "build_variants", f"lib_{lib}.so")
    env = dict(os.environ)
    if lib != "tree":
        env["ILCC2_LIB"] = path
    for a in assign:
        k, v = a.split("=", 1)
        env[k] = v
    subprocess.run([sys.executable, __file__, "--child", name, str(F)], env=env)
