#!/bin/bash
# compute-sanitizer over one pass of the hot path (device path) and one host-path call (staging rings), small batches
OUT=${SAN_OUT:-gpurun_out/sanitizer_r2.txt}
: > $OUT
for F in 40 400; do
  echo "== memcheck, $F frames (device path)" >> $OUT
  timeout 600 compute-sanitizer --tool memcheck python tools/profile_step.py 1 $F 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|ok " | head -5 >> $OUT
done
echo "== memcheck, host path 200 frames pageable 32-byte records (staging rings + repack)" >> $OUT
timeout 600 compute-sanitizer --tool memcheck python - >> $OUT 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from lidar_camera_calibration_b200 import api, synth
F = 200
pts, clicks = synth.make_batch(F)
rec = np.zeros((F, pts.shape[1], 8), np.float32); rec[:, :, :3] = pts[:, :, :3]; rec[:, :, 4] = pts[:, :, 3]
ex = api.CornerExtractor(F, synth.VLP16.n_points, roi_capacity=8192)
st, c = ex.extract(rec, clicks)
st2, c2 = ex.extract(pts, clicks)
print("host path ok", int((st == 0).sum()), "equal", bool(np.array_equal(c, c2, equal_nan=True)))
PY
grep -c . $OUT > /dev/null
echo "== racecheck, 24 frames (device path)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/profile_step.py 1 24 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|ERROR|WARN" | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20 >> $OUT
tail -40 $OUT
