"""Per-unit cycle counters of K5 (library built with -DFIT_TIMING): python tools/fit_timing.py 1024"""
import os, sys
os.environ.setdefault('ILCC2_E2E_CHUNK', '1000000')  # one launch for the whole batch, as in the device-resident path
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
print(' job duration (hypothesis-0 CTA, warp 0), kilo-cycles: mean %.0f  p50 %.0f  p90 %.0f  max %.0f;  sum over frames / 1e6 = %.1f Mcycles' % (tot.mean(), np.median(tot), np.quantile(tot, 0.9), tot.max(), tot.sum() / 1e3))
n_opt = np.array([x['n_opt'] for x in d]); order = np.argsort(-n_opt)
for lo, hi in ((0, 50), (50, 103), (103, 200), (200, 500), (500, 1024)):
    sel = order[lo:hi]
    print('  frames ranked %4d..%4d by n_opt (n %4d..%4d): job kc mean %.0f max %.0f  [classify %.0f lists %.0f scalar %.0f]' % (lo, hi, n_opt[sel].max(), n_opt[sel].min(), tot[sel].mean(), tot[sel].max(), pl[sel, 0].mean(), pl[sel, 1:3].sum(1).mean(), pl[sel, 3].mean()))
