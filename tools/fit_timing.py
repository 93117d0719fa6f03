import sys
sys.path.insert(0, '.')
import numpy as np
from lidar_camera_calibration_b200 import api, synth
F = int(sys.argv[1])
pts, clicks = synth.make_batch(F)
ex = api.CornerExtractor(F, synth.VLP16.n_points)
for _ in range(2):
    st, c, d = ex.extract(pts, clicks, want_diag=True)
pl = np.array([x['plane'] for x in d])
names = ['classify', 'list_cost', 'list_jac', 'scalar']
print('F', F, 'hypothesis-0 warp, kilo-cycles:', {n: (round(float(pl[:, i].mean())), round(float(pl[:, i].max()))) for i, n in enumerate(names)})
tot = pl.sum(1); i = int(tot.argmax())
print(' slowest frame', i, pl[i].round(), 'n_opt', d[i]['n_opt'], 'evals', d[i]['n_evals'], 'cost', d[i]['cost'], 'hyp', d[i]['hyp_chosen'])
