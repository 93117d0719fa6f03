import sys, time
sys.path.insert(0, '.')
import numpy as np
from lidar_camera_calibration_b200 import api, synth
start, count = int(sys.argv[1]), int(sys.argv[2])
pts, clicks = synth.make_batch(count, start=start)
ex = api.CornerExtractor(count, synth.VLP16.n_points)
ex.set_profiling(True)
t = time.time()
st, corners, diags = ex.extract(pts, clicks, flags=api.FLAG_AUTO, want_diag=True)
print(start, count, 'ok', int((st == 0).sum()), 'time %.3f' % (time.time() - t), ex.stage_times(),
      'max roi', max(d['n_roi'] for d in diags), 'max evals', max(int(d['n_evals'].max()) for d in diags), flush=True)
