"""Multi-GPU from a C++ host (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

host/multi_gpu_corners.cpp forks one process per GPU, builds an NCCL communicator and drives the C ABI only:
ilcc2_create_sized, ilcc2_gather_setup_nccl, ilcc2_extract_corners(ILCC2_MEM_DEVICE), ilcc2_gather_sync.  Every rank
must hold, after each step, exactly what a single GPU computes for all frames of that step (frames are independent:
SURVEY.md 8(e) "all-gathered corners are byte-identical to the 1-GPU run"), with consecutive steps producing different
corners in every slot (the double buffer of the fused gather)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2])
def test_cpp_host_gathers_every_ranks_corners(product_lib, tmp_path, world):
    from lidar_camera_calibration_b200 import api, synth
    N, K = 48, 3
    pts, clicks = synth.make_batch(N, start=700)
    pts.tofile(tmp_path / "frames.bin")
    clicks.astype(np.float32).tofile(tmp_path / "clicks.bin")
    tool = os.path.join(ROOT, "lidar_camera_calibration_b200", "host", "multi_gpu_corners")
    r = subprocess.run([tool, "--frames", str(tmp_path / "frames.bin"), "--clicks", str(tmp_path / "clicks.bin"),
                        "--n-frames", str(N), "--points", str(synth.VLP16.n_points), "--gpus", str(world), "--steps", str(K),
                        "--out", str(tmp_path / "g")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout.strip())
    st1, c1 = api.CornerExtractor(N, synth.VLP16.n_points).extract(pts, clicks, flags=api.FLAG_AUTO)
    per = N // world
    for step in range(K):
        # every shard is rotated by `step` frames
        order = np.concatenate([r0 * per + (np.arange(per) + step) % per for r0 in range(world)])
        want_c, want_s = c1[order], st1[order]
        for rank in range(world):
            raw = np.fromfile(tmp_path / f"g.rank{rank}.step{step}.bin", dtype=np.uint8)
            got_c = raw[: N * 35 * 12].view(np.float32).reshape(N, 35, 3)
            got_s = raw[N * 35 * 12:].view(np.int32)
            assert np.array_equal(got_s, want_s), (step, rank)
            assert np.array_equal(got_c, want_c, equal_nan=True), (step, rank)


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("parts,multicast", [(2, 1), (3, 1), (3, 0)])
def test_fused_gather_lanes(product_lib, parts, multicast):
    """dist.FusedGather under pipelined lanes on two GPUs (torchrun): inline barrier (2 parts) and the barrier on a side
    stream one step behind the lane (3 parts), through the NVLS multicast mapping and with peer stores; seven steps on
    different frames, every step's gathered corners equal to a plain NCCL all-gather on every rank."""
    import sys
    worker = os.path.join(ROOT, "tests", "helpers", "fused_gather_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29530 + parts * 2 + multicast),
                        worker, str(parts), str(multicast)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-600:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ok True" in r.stdout
