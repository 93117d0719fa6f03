"""CPU analysis (oracle only): how far do consecutive evaluation points of the board-pose solves move?  Used to size
the incremental ("slack") classification of K5.  python tools/trace_solver.py [frames]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from lidar_camera_calibration_b200 import synth
L = O.lib()
L.ilcc2o_trace.argtypes = [C.c_void_p, C.c_int]; L.ilcc2o_trace_count.restype = C.c_int
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pts, clicks = synth.make_batch(F)
p = O.default_params()
W, H, g = p.cells_x, p.cells_y, p.grid_length
rows = []
for f in range(F):
    st, corners, d, roi, clu, brd = O.extract_frame(pts[f], clicks[f], flags=O.FLAG_AUTO, want_indices=True)
    if st != 0:
        continue
    P = pts[f][brd]
    T = np.asarray(d["pca"], np.float64).reshape(4, 4)
    q = P[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    I = P[:, 3].astype(np.float64)
    keep = (I < d["gray"][0]) | (I > d["gray"][1])
    yz = np.ascontiguousarray(q[keep][:, 1:3]); white = np.ascontiguousarray((I[keep] > d["gray"][1]).astype(np.uint8))
    rmax = np.abs(yz[:, 0]).max() + np.abs(yz[:, 1]).max()
    for tlw in (0, 1):
        x = np.zeros(3)
        for oob in (1, 0):
            buf = np.zeros((400, 4)); L.ilcc2o_trace(buf.ctypes.data, 400)
            term, x, cost, it, ne = O.solve(yz, white, x, W, H, g, bool(tlw), bool(oob))
            n = L.ilcc2o_trace_count(); L.ilcc2o_trace(None, 0)
            tr = buf[:n]
            rows.append((f, tlw, oob, tr, rmax, yz, white, cost, it))
print("solves", len(rows))
# policy simulation: ref = pose of the last full classification; delta in cells; selected fraction ~ points with slack <= delta
def slack_cells(yz, x, W, H, g):
    c, s = np.cos(x[0]), np.sin(x[0])
    fi = ((yz[:, 0] * c - yz[:, 1] * s) + x[1] + W * g / 2) / g
    fj = ((yz[:, 1] * c + yz[:, 0] * s) + x[2] + H * g / 2) / g
    di = np.abs(fi - np.rint(fi)); dj = np.abs(fj - np.rint(fj))
    edge = np.maximum(np.abs(fi - W / 2) - W / 2, np.abs(fj - H / 2) - H / 2)
    return np.minimum(np.minimum(di, dj), np.abs(edge)), fi, fj, edge
tot_full = tot_cull = tot_evalpts = 0.0
hist = []
for (f, tlw, oob, tr, rmax, yz, white, cost, it) in rows:
    n = len(yz); ref = None
    for e in range(len(tr)):
        x = tr[e, :3]
        if tr[e, 3] == 1 and e > 0:      # Jacobian at the accepted point = same pose as the previous evaluation
            continue
        if ref is None:
            ref = x.copy(); sl = slack_cells(yz, ref, W, H, g)[0]; tot_full += n; continue
        delta = (rmax * abs(x[0] - ref[0]) + max(abs(x[1] - ref[1]), abs(x[2] - ref[2]))) / g
        sel = (sl <= delta + 1e-4).mean()
        hist.append((delta, sel))
        if sel > 0.4:
            ref = x.copy(); sl = slack_cells(yz, ref, W, H, g)[0]; tot_full += n
        else:
            tot_cull += n; tot_evalpts += sel * n
h = np.array(hist)
print("evaluations (excl. first):", len(h), " median delta[cells] %.4f  p90 %.4f" % (np.median(h[:, 0]), np.percentile(h[:, 0], 90)))
print("selected fraction: median %.3f mean %.3f; share of evals with sel>0.4: %.3f" % (np.median(h[:, 1]), h[:, 1].mean(), (h[:, 1] > 0.4).mean()))
base = sum(len(r[5]) * ((r[3][:, 3] == 0).sum() + 1) for r in rows)
print("point-classifications: baseline %.3g  policy: full %.3g + cull-scan %.3g (cheap) + extra exact evals %.3g" % (base, tot_full, tot_cull, tot_evalpts))
