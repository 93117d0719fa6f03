#!/bin/bash
# compute-sanitizer over the launch shapes of the throughput mode (ilcc2_set_lanes -> K5 with one warp per solve)
OUT=gpurun_out/sanitizer_r3.txt
: > $OUT
for F in 40 400; do
  echo "== memcheck, $F frames, lanes 64 (device path, board_fit_kernel<1,1>)" >> $OUT
  timeout 600 compute-sanitizer --tool memcheck python tools/profile_step.py 1 $F 64 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|ok " | head -5 >> $OUT
done
echo "== racecheck, 24 frames, lanes 64 (device path, board_fit_kernel<1,1>)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/profile_step.py 1 24 64 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|ERROR|WARN|ok " | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20 >> $OUT
echo "== memcheck, pipelined lanes: 3 handles x 3 streams, 5 batches of 96 frames + host lanes" >> $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k pipelined 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed" | head -5 >> $OUT
cat $OUT
