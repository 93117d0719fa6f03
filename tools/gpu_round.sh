#!/bin/bash
# One GPU-box visit: bench lines of both arms, the ncu launch list of the bench command, one ncu --set full capture of
# a step (all kernels).  usage: bash tools/gpu_round.sh <tag>
TAG=${1:-r2x}
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -s 12 -c 6 -o gpurun_out/ncu_all_$TAG python tools/profile_step.py 3 > gpurun_out/ncu_all_$TAG.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','single_frame_latency_ms','scratch_mb')})
print('e2e', d['e2e']['value'], {k:(round(v['ms_per_step'],2) if 'ms_per_step' in v else v) for k,v in d['e2e_legs'].items()})
print('parity', d['parity']); print('clocks', d['clocks']); print('hdl64', d['hdl64']); print('acc', d['accuracy_vs_truth_m'])
print('cpu', d.get('cpu_baseline'))
r=json.load(open('gpurun_out/bench_ref_$TAG.json')); print('ref', r['value'], r['cpu_baseline'])
"
