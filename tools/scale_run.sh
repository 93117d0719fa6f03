#!/bin/bash
# bash tools/scale_run.sh <N> <tag> [extra bench args]: the driver's launch line for N ranks on one node
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
lines=[l for l in open("gpurun_out/bench_${N}gpu_$TAG.json") if l.startswith("{")]
d=json.loads(lines[-1])
print("N=$N", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],2), d["config"]["collective"], d["config"]["gather_equals_nccl_all_gather"], d["config"].get("numa_binding"))
print("  legs", {k:(round(v["ms_per_step"],2) if "ms_per_step" in v else v) for k,v in d["e2e_legs"].items()})
PY
