"""torchrun --nproc-per-node N tools/band_check.py : one 4K frame pair split by block-row
bands over N GPUs, field gathered with one NCCL all_gather per array, compared on every rank
with the unsharded search.  Prints the device time of the banded search (max over ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402
from motionestimation_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, B, R = 3840, 2160, 16, int(os.environ.get("BAND_R", "64"))
    cur8, ref8 = me.tiled_frames(W, H)
    cur, ref = torch.from_numpy(cur8).cuda(), torch.from_numpy(ref8).cuda()
    with me.Estimator(W, H, B, R, device=local) as est:
        full = est.search_u8(cur8, ref8)
        # unsharded reference time on this rank (device-resident)
        nb = est.num_blocks
        o = [torch.zeros((1, nb), dtype=torch.int32, device="cuda") for _ in range(3)]
        for _ in range(3):
            est.search_device(cur, ref, W, W * H, 1, o[0], o[1], o[2])
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        est.search_device(cur, ref, W, W * H, 1, o[0], o[1], o[2])
        s1.record()
        torch.cuda.synchronize()
        t_single = s0.elapsed_time(s1)
        for _ in range(3):
            res = sharding.search_banded(est, cur, ref, W, W * H, 1)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sharding.search_banded(est, cur, ref, W, W * H, 1)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = (np.array_equal(res["mvx"].cpu().numpy(), full["mvx"]) and
              np.array_equal(res["mvy"].cpu().numpy(), full["mvy"]) and
              np.array_equal(res["ssd"].cpu().numpy().view(np.uint32), full["ssd"]) and
              np.array_equal(res["score"].cpu().numpy().view(np.uint32), full["score"].view(np.uint32)))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"band sharding x{world}: 4K 16x16 +-{R} one pair, banded search + gather {t.item():.3f} ms "
                  f"(max over ranks) vs {t_single:.3f} ms unsharded on one GPU, identical on all ranks: "
                  f"{bool(flag.item())}", flush=True)
        assert ok
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
