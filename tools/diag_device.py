import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from lidar_camera_calibration_b200 import api, synth
F = int(sys.argv[1]); mode = sys.argv[2]
pts, clicks = synth.make_batch(min(F, 128))
reps = (F + pts.shape[0] - 1) // pts.shape[0]
pts = np.ascontiguousarray(np.tile(pts, (reps, 1, 1))[:F]); clicks = np.ascontiguousarray(np.tile(clicks, (reps, 1))[:F])
ex = api.CornerExtractor(F, synth.VLP16.n_points)
t = time.time()
if mode == 'host':
    st, c = ex.extract(pts, clicks, flags=api.FLAG_AUTO)
else:
    dp, dc = torch.from_numpy(pts).cuda(), torch.from_numpy(clicks).cuda()
    st, c = ex.extract_torch(dp, dc, flags=api.FLAG_AUTO); torch.cuda.synchronize(); st = st.cpu().numpy()
print('F', F, mode, 'ok', int((st == 0).sum()), 'time %.3f' % (time.time() - t), flush=True)
