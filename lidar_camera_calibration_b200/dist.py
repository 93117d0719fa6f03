"""Multi-GPU plumbing: one process per GPU, frames sharded in contiguous blocks, one exchange step.

Frames are independent units, so the data path needs no collective; the only exchange is the
all-gather of the per-frame corners (+ status) before the host-side PnP (ilcc2/test/calib_lidar_cam.cpp:116-165
consumes all frames' corners).  torch.distributed carries it: NCCL over NVLink on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import os


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world_size, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int):
    """Pin this process (and the threads / page-locked buffers it creates afterwards) to the CPUs of the NUMA node its
    GPU hangs off, read from sysfs (/sys/bus/pci/devices/<bus id>/local_cpulist).  With one rank per GPU the host
    buffers of every rank otherwise tend to land on one socket, and H2D copies of the other socket's GPUs cross the
    inter-socket link.  Returns a short description, or None when nothing was changed."""
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed or allowed == set(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        node = open(f"{base}/numa_node").read().strip()
        return f"gpu {local_rank} ({bus}): numa node {node}, {len(allowed)} cpus"
    except Exception:  # noqa: BLE001
        return None


def shard_range(n_frames: int, world_size: int, rank: int):
    """Contiguous block of frames owned by `rank` (same arithmetic as ilcc2_shard_range in the C ABI)."""
    b = n_frames * rank // world_size
    e = n_frames * (rank + 1) // world_size
    return b, e - b


def allgather_corners(corners, status, world_size: int, out_corners=None, out_status=None):
    """corners (F_local, ncorn, 3) float32, status (F_local,) int32 -- equal F_local on every rank.
    Returns (all_corners (world*F_local, ncorn, 3), all_status (world*F_local,)) on every rank."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return corners, status
    F = corners.shape[0]
    if out_corners is None:
        out_corners = torch.empty((world_size * F,) + tuple(corners.shape[1:]), dtype=corners.dtype, device=corners.device)
    if out_status is None:
        out_status = torch.empty((world_size * F,), dtype=status.dtype, device=status.device)
    dist.all_gather_into_tensor(out_corners, corners.contiguous())
    dist.all_gather_into_tensor(out_status, status.contiguous())
    return out_corners, out_status


def allgather_rows(x, world_size: int):
    """Concatenation over ranks (rank order) of a tensor with the same shape on every rank."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return x
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out


class FusedGather:
    """The all-gather of the per-frame corners fused into the kernel that produces them.

    Every rank owns a symmetric-memory buffer of `halves` parts [halves][world * F_local frames]
    (torch.distributed._symmetric_memory: the allocations of all ranks are mapped into each other's address space over
    NVLink); K5 stores frame f of rank r into slot r * F_local + f of EVERY rank's current part
    (ilcc2_set_gather_targets) -- through the NVLS multicast mapping with one multimem.st per value when the fabric
    offers it, else with one plain store per peer.  barrier() is the only cross-rank operation of a step: a signal-pad
    barrier, no data moves through a collective.

    halves = 2 (default): barrier() runs on the current stream and flips to the other part for the next step, so a rank
    that runs ahead cannot overwrite corners another rank is still reading: it reaches this part again only after the
    NEXT barrier, which every rank enters after it has consumed this step (consume(k) precedes produce(k+1) in stream
    order).  `corners` / `status` name the part the last barrier() completed.

    halves = 3 (deferred): barrier() only records the producer on the current stream; the signal-pad barrier itself runs
    on a side stream, so the ranks' skew is absorbed while the current stream already works on its next step.  Step k
    writes part k mod 3; produce(k+3) overwrites it and needs every rank's consume(k) to be over: consume(k) is
    enqueued behind barrier(k) (wait_ready) and therefore ahead of barrier(k+1) on each rank's side stream, so the
    current stream waits for ITS barrier(k+1) before produce(k+3) -- i.e. barrier() of step k ends with a wait for the
    barrier of step k-1, which guards the next produce.  A consumer calls wait_ready(stream), reads `corners` / `status`
    (the part of the last step handed to barrier()) on that stream, and calls release(stream), which puts the read
    ahead of the next barrier on the side stream.

    Raises if symmetric memory cannot be set up; callers fall back to allgather_corners (NCCL)."""

    def __init__(self, extractor, frames_local: int, n_corners: int, rank: int, world: int, device,
                 use_multicast: bool = True, halves: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        assert halves in (2, 3)
        self.world, self.rank, self.F, self.ncorn, self.halves = world, rank, frames_local, n_corners, halves
        total = world * frames_local
        cbytes = total * n_corners * 3 * 4
        half = (cbytes + total * 4 + 255) // 256 * 256
        self.buf = symm.empty((halves * half,), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self._views = []
        for k in range(halves):
            b = self.buf[k * half:(k + 1) * half]
            self._views.append((b[:cbytes].view(torch.float32).view(total, n_corners, 3),
                                b[cbytes:cbytes + total * 4].view(torch.int32)))
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        if mc:
            self.mode, self._kind = "nvls multicast, multimem.st", 1
            self._targets = [([mc + k * half], [mc + k * half + cbytes]) for k in range(halves)]
        else:
            self.mode, self._kind = "peer stores", 0
            bases = [int(p) for p in self.hdl.buffer_ptrs]
            self._targets = [([p + k * half for p in bases], [p + k * half + cbytes for p in bases]) for k in range(halves)]
        self._extractor, self._slot, self._half = extractor, rank * frames_local, 0
        self.corners, self.status = self._views[0]
        self._point(0)
        self.hdl.barrier(channel=0)
        if halves == 3:
            self._side = torch.cuda.Stream(device)
            self._produced = torch.cuda.Event()
            self._bar_events = [torch.cuda.Event(), torch.cuda.Event()]   # barrier(k) -> slot k & 1
            self._steps = 0
            torch.cuda.current_stream(device).synchronize()

    def _point(self, k: int):
        c, s = self._targets[k]
        self._extractor.set_gather_targets(c, s, self._slot, self._kind)

    def barrier(self):
        """halves = 2: all ranks' stores of this step are visible in every rank's buffer afterwards (runs on the current
        stream); the next step's stores go to the other part.  halves = 3: the same barrier, on the side stream (see the
        class comment); the current stream only waits for the barrier of the PREVIOUS step."""
        import torch
        if self.halves == 2:
            self.hdl.barrier(channel=0)
            self.corners, self.status = self._views[self._half]
            self._half ^= 1
            self._point(self._half)
            return
        cur = torch.cuda.current_stream(self.buf.device)
        k = self._steps
        self._produced.record(cur)
        self._side.wait_event(self._produced)
        with torch.cuda.stream(self._side):
            self.hdl.barrier(channel=0)
        self._bar_events[k & 1].record(self._side)
        if k >= 1:
            cur.wait_event(self._bar_events[(k - 1) & 1])   # guards produce(k+1) ... produce(k+2) on this stream
        self.corners, self.status = self._views[self._half]
        self._half = (self._half + 1) % 3
        self._point(self._half)
        self._steps = k + 1

    def wait_ready(self, stream=None):
        """Make `stream` (default: current) wait until `corners` / `status` hold every rank's results of the last step
        given to barrier().  A no-op with halves = 2 when the consumer runs on the stream barrier() ran on."""
        import torch
        if self.halves == 3 and self._steps > 0:
            s = stream if stream is not None else torch.cuda.current_stream(self.buf.device)
            s.wait_event(self._bar_events[(self._steps - 1) & 1])

    def release(self, stream=None):
        """The reads of `corners` / `status` enqueued on `stream` so far come before this rank's next barrier (halves = 3;
        with halves = 2 stream order already says so when the consumer runs on the stream of barrier())."""
        import torch
        if self.halves == 3:
            s = stream if stream is not None else torch.cuda.current_stream(self.buf.device)
            ev = torch.cuda.Event()
            ev.record(s)
            self._side.wait_event(ev)

    def close(self):
        self._extractor.set_gather_targets([], [], 0)
