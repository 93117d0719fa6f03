"""CPU, world_size 2 over gloo: frame sharding + the all-gather of per-frame corners.
The compute of each rank is stood in for by the CPU oracle (test infrastructure); what is under test is the host
logic: contiguous shards, gathered result byte-identical to the single-process run."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import torch.distributed as dist
from lidar_camera_calibration_b200 import dist as ldist, synth
from oracle import oracle as O
rank, world, _ = ldist.init_from_env("gloo")
F = 6
b, c = ldist.shard_range(F, world, rank)
pts, clicks = synth.make_batch(c, start=b)
st, corners = O.extract_batch(pts, clicks, flags=O.FLAG_AUTO, threads=1)
all_c, all_s = ldist.allgather_corners(torch.from_numpy(corners), torch.from_numpy(st), world)
uv = torch.from_numpy(synth.project_truth(synth.make_batch(c, start=b, want_truth=True)[2], start=b))
all_uv = ldist.allgather_rows(uv, world)
if rank == 0:
    np.save({out!r} + "_corners.npy", all_c.numpy()); np.save({out!r} + "_status.npy", all_s.numpy())
    np.save({out!r} + "_uv.npy", all_uv.numpy())
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_gather_equals_single_process(tmp_path, oracle):
    from lidar_camera_calibration_b200 import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    pts, clicks = synth.make_batch(6)
    st, corners = oracle.extract_batch(pts, clicks, flags=oracle.FLAG_AUTO, threads=2)
    assert np.array_equal(np.load(out + "_status.npy"), st)
    assert np.array_equal(np.load(out + "_corners.npy"), corners, equal_nan=True)
    # the camera-side rows of configs[4] (projected ground truth, per-frame Philox noise) gather in rank order too
    gt = synth.make_batch(6, want_truth=True)[2]
    assert np.array_equal(np.load(out + "_uv.npy"), synth.project_truth(gt))
