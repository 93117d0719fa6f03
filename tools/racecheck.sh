#!/bin/bash
# racecheck of one pass over 24 frames; per severity the distinct (write site, read/write site) pairs
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200000 python tools/profile_step.py 1 24 $1 > /tmp/race.log 2>&1
grep -E "RACECHECK SUMMARY" /tmp/race.log
python - <<'PY'
import re, collections
agg = collections.Counter()
lines = open("/tmp/race.log", errors="replace").read().split("\n")
i = 0
while i < len(lines):
    m = re.match(r"========= (Error|Warning): (\(Warp Level Programming\) )?Potential (\w+) hazard detected at (__shared__|__global__)", lines[i])
    if m and i + 2 < len(lines):
        def site(l):
            f = re.search(r"at (?:\[clone .*?\] )?(?:\w+ )?(?:ilcc2::)?(\w+)", l)
            ln = re.search(r"(\w+\.(?:cu|hpp|h)):(\d+)", l)
            return (f.group(1) if f else "?") + "@" + (ln.group(1) + ":" + ln.group(2) if ln else "?")
        agg[(m.group(1), "warp-level" if m.group(2) else "block-level", m.group(3), site(lines[i + 1]), site(lines[i + 2]))] += 1
        i += 3
    else:
        i += 1
for k, v in sorted(agg.items(), key=lambda kv: (kv[0][0], -kv[1])):
    print(v, *k)
PY
