"""torchrun --nproc-per-node N tools/band_check.py : one 4K frame pair split by block-row
bands over N GPUs; the field is completed (a) with one NCCL all_gather of the packed arrays and
(b) with no collective at all: peer-mapped fields, the search kernel stores into every rank's copy,
one device-side flag barrier.  Both are compared on every rank with the unsharded search.  Prints
the device times (max over ranks)."""
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
        tn = []
        for _ in range(5):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = sharding.search_banded(est, cur, ref, W, W * H, 1)
            e1.record()
            torch.cuda.synchronize()
            tn.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(tn)[len(tn) // 2]], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = (np.array_equal(res["mvx"].cpu().numpy(), full["mvx"]) and
              np.array_equal(res["mvy"].cpu().numpy(), full["mvy"]) and
              np.array_equal(res["ssd"].cpu().numpy().view(np.uint32), full["ssd"]) and
              np.array_equal(res["score"].cpu().numpy().view(np.uint32), full["score"].view(np.uint32)))
        # (b) peer-mapped fields: no collective
        field = sharding.PeerField(est, 1)
        for _ in range(3):
            resp = sharding.search_banded_peer(est, field, cur, ref, W, W * H, 1, check=False)
        torch.cuda.synchronize()
        dist.barrier()
        tp = []
        for _ in range(5):
            dist.barrier()
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            resp = sharding.search_banded_peer(est, field, cur, ref, W, W * H, 1, check=False)
            p1.record()
            torch.cuda.synchronize()
            tp.append(p0.elapsed_time(p1))
        tpeer = torch.tensor([sorted(tp)[len(tp) // 2]], device="cuda")
        dist.all_reduce(tpeer, op=dist.ReduceOp.MAX)
        timed_out = est.peer_barrier_timed_out()
        okp = (not timed_out and np.array_equal(resp["mvx"].cpu().numpy(), full["mvx"]) and
               np.array_equal(resp["mvy"].cpu().numpy(), full["mvy"]) and
               np.array_equal(resp["ssd"].cpu().numpy().view(np.uint32), full["ssd"]) and
               np.array_equal(resp["score"].cpu().numpy().view(np.uint32), full["score"].view(np.uint32)))
        del resp
        torch.cuda.synchronize()
        field.close()
        flag = torch.tensor([1 if ok else 0, 1 if okp else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"band sharding x{world}: 4K 16x16 +-{R} one pair: unsharded {t_single:.3f} ms | banded search + "
                  f"NCCL all_gather {t.item():.3f} ms, identical on all ranks: {bool(flag[0].item())} | banded search "
                  f"storing into peer-mapped fields + device flag barrier {tpeer.item():.3f} ms, identical on all "
                  f"ranks: {bool(flag[1].item())} (max over ranks)", flush=True)
        assert ok and okp
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
