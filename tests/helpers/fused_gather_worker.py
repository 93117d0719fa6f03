"""Worker of tests/test_gpu_multi.py::test_fused_gather_lanes (one process per GPU under torchrun): a stream of DIFFERENT
batches through pipelined lanes, every lane with its own dist.FusedGather; after each step every rank must hold what a
plain NCCL all-gather of that step's corners / status delivers.  argv: parts (2 | 3), multicast (0 | 1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lidar_camera_calibration_b200 import api, dist as ldist, synth  # noqa: E402

parts, multicast = int(sys.argv[1]), bool(int(sys.argv[2]))
rank, world, local = ldist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
F, L, STEPS = 40, 2, 7
N = synth.VLP16.n_points
pipe = api.PipelinedExtractor(L, F, N, device=local, roi_capacity=8192)
fused = [ldist.FusedGather(pipe.extractors[j], F, pipe.n_corners, rank, world, dev, use_multicast=multicast, halves=parts)
         for j in range(L)]
batches = []
for k in range(STEPS):   # every rank and step its own frames
    p, c = synth.make_batch(F, start=5000 + 977 * k + 131 * rank)
    batches.append((torch.from_numpy(p).to(dev), torch.from_numpy(c).to(dev)))
outs = [(torch.empty((F, pipe.n_corners, 3), dtype=torch.float32, device=dev), torch.empty((F,), dtype=torch.int32, device=dev))
        for _ in range(STEPS)]
got = []
cur = torch.cuda.current_stream(dev)
for k, (p, c) in enumerate(batches):
    j = pipe.submit(p, c, outs[k][0], outs[k][1], after=lambda lane: fused[lane].barrier())
    # the consumer of step k: a copy of the gathered part, on the lane's stream behind the exchange
    with torch.cuda.stream(pipe.streams[j]):
        fused[j].wait_ready()
        got.append((fused[j].corners.clone(), fused[j].status.clone()))
        fused[j].release()
pipe.join(cur)
for fg in fused:
    fg.wait_ready(cur)
torch.cuda.synchronize(dev)
ok = True
for k in range(STEPS):
    ref_c, ref_s = ldist.allgather_corners(outs[k][0], outs[k][1], world)
    same = torch.equal(ref_s, got[k][1]) and torch.equal(torch.nan_to_num(ref_c, nan=-1.0), torch.nan_to_num(got[k][0], nan=-1.0))
    distinct = k == 0 or not torch.equal(torch.nan_to_num(got[k][0], nan=-1.0), torch.nan_to_num(got[k - 1][0], nan=-1.0))
    ok = ok and same and distinct
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"fused gather lanes: parts {parts} mode '{fused[0].mode}' steps {STEPS} ok {bool(t.item())}", flush=True)
for fg in fused:
    fg.close()
pipe.close()
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
