"""torchrun --nproc-per-node N tools/h2d_probe.py : copy-only host->device bandwidth of the box.

Names the limiter of the end-to-end arm at 8 GPUs (e2e 6.34x of 8): every rank owns one
GPU and a pinned host buffer; for a list of rank SUBSETS the active ranks copy the buffer to their GPU
back to back (cudaMemcpyAsync through torch, CUDA events) while the others idle.  If two GPUs share a
PCIe uplink the pair {a, b} drops while {a, c} does not; if the host memory system is the limit, only
the total matters.  Variants: pinned memory allocated with / without a per-rank CPU affinity (NUMA
first touch), write-combined pinned memory, and the copy size the bench uses (one slot = 16 pairs of
1080p = 2 x 33 MB).  Rank 0 prints one JSON object; the table goes to profiles/.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = me.load_library()
    all_cpus = sorted(os.sched_getaffinity(0))
    nbytes = 64 << 20
    reps = 24
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def host_buf(wc):
        ptr = lib.me_b200_host_alloc_ex(nbytes, me.ME_HOST_WRITE_COMBINED if wc else 0)
        a = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr))
        a[:] = rank + 1          # first touch by this rank's thread
        return ptr, torch.from_numpy(a)

    def measure(active, host, size=nbytes):
        """GB/s of this rank (0 when idle) with exactly the ranks in `active` copying."""
        dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank in active:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                dev[:size].copy_(host[:size], non_blocking=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                dev[:size].copy_(host[:size], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            gbs = size * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(x.item()), 2) for x in out]

    subsets = [[0]]
    if world >= 2:
        subsets += [[0, 1]]
    if world >= 4:
        subsets += [[0, 2], [0, 3], [0, 1, 2, 3]]
    if world >= 8:
        subsets += [[0, 4], [0, 7], [4, 5, 6, 7], [0, 2, 4, 6], [1, 3, 5, 7], list(range(8))]
    res = {"world": world, "host_cpus": len(all_cpus), "copy_bytes": nbytes, "reps": reps, "unit": "GB/s per rank",
           "runs": []}

    variants = [("pinned, no affinity", False, False), ("pinned, per-rank affinity", True, False),
                ("write-combined pinned, per-rank affinity", True, True)]
    for label, aff, wc in variants:
        if aff and len(all_cpus) >= 2 * world:
            per = len(all_cpus) // world
            os.sched_setaffinity(0, all_cpus[local * per:(local + 1) * per])
        else:
            os.sched_setaffinity(0, all_cpus)
        ptr, host = host_buf(wc)
        for sub in subsets:
            g = measure(sub, host)
            if rank == 0:
                res["runs"].append({"variant": label, "active": sub, "gbs": g, "total": round(sum(g), 1)})
        # the size the bench's slot upload uses (16 pairs of 1080p, one frame array = 33 MB)
        g = measure(list(range(world)), host, 16 * 1920 * 1080)
        if rank == 0:
            res["runs"].append({"variant": label + ", 33 MB copies", "active": list(range(world)), "gbs": g,
                                "total": round(sum(g), 1)})
        del host
        lib.me_b200_host_free(ptr)
    os.sched_setaffinity(0, all_cpus)

    # host memory bandwidth with all ranks reading at once (numpy copy of 256 MB, pageable)
    src = np.ones(256 << 20, np.uint8)
    dst = np.empty_like(src)
    dist.barrier()
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        np.copyto(dst, src)
    dt = time.perf_counter() - t0
    t = torch.tensor([2 * 3 * src.nbytes / dt / 1e9], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank == 0:
        res["host_memcpy_gbs_read_plus_write_all_ranks_at_once"] = [round(float(x.item()), 1) for x in out]
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        except Exception as ex:  # pragma: no cover
            res["topo"] = repr(ex)
        try:
            res["lscpu"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines()
                            if any(k in l for k in ("Model name", "Socket", "NUMA", "CPU(s):", "Thread"))]
        except Exception as ex:  # pragma: no cover
            res["lscpu"] = repr(ex)
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
