#!/usr/bin/env python
"""bench.py -- headline benchmark of the full-search path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

A STEP is one pass of the hot path (the replacement of src/cpu/main.c:144-158)
over one batch of synthetic frame pairs.  Default workload = BASELINE.json
configs[1]: 1920x1080 luma, 16x16 blocks, full search +-32 (Beauty is absent from
the reference checkout, so the frames are the seeded synthetic stand-in of
SURVEY.md section 8d).  One process per GPU (torchrun for N > 1), pairs sharded by
rank, no data-path collective (frame pairs are independent) => weak scaling.

Printed JSON line (rank 0):
  value     frames/s, whole job, inputs resident in HBM, CUDA events on the launch stream,
            barrier + synchronize on both sides, max over ranks
  e2e       same metric through the host-buffer C ABI (me_b200_submit / me_b200_wait):
            pinned host frames -> H2D -> search -> D2H of the motion field, every step; the K
            steps are streamed through the 4-slot ring (batch i+1 uploads while batch i is
            searched); "step_synchronous" drains the ring after every step instead
  roofline  integer-pipe roofline of the search kernel: algorithmic lane-instructions
            (0.5 per pixel-compare: VABSDIFF4 + IDP.4A per 4 pixels) / step time,
            against the VABSDIFF4+IDP.4A pair rate measured live by me_b200_int_peak
  cpu_baseline  the unmodified reference CPU path (oracle/_ref) on this box's host cores
--impl reference times the reference's own CPU implementation instead (no GPU work).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, B, R, pairs per GPU per step, description)
    "1080p_16x16_pm32": (1920, 1080, 16, 32, 64, "BASELINE configs[1]: synthetic 1920x1080 luma, 16x16 blocks, full search +-32"),
    "1080p_16x16_pm64": (1920, 1080, 16, 64, 32, "BASELINE configs[2]: synthetic 1920x1080 luma, 16x16 blocks, full search +-64"),
    "4k_8x8_pm32": (3840, 2160, 8, 32, 16, "BASELINE configs[4]: synthetic 3840x2160 luma, 8x8 blocks, full search +-32"),
    "4k_16x16_pm32": (3840, 2160, 16, 32, 16, "BASELINE configs[4]: synthetic 3840x2160 luma, 16x16 blocks, full search +-32"),
    "foreman_8x8_pm12": (352, 288, 8, 12, 512, "BASELINE configs[0]: Foreman YF2->YF1, reference defaults 8x8 +-12"),
    # the only configuration the reference publishes numbers for: 3840x2160, 8x8, +-12
    "4k_8x8_pm12": (3840, 2160, 8, 12, 16, "reference's own published runs: synthetic 3840x2160 luma, 8x8 blocks, full search +-12"),
    # the memory-bound end (north_star's "memory-bound small-range cases"): the
    # only candidate is the co-located block, one streaming pass over both frames
    "1080p_16x16_pm0": (1920, 1080, 16, 0, 256, "memory-bound small-range case: synthetic 1920x1080 luma, 16x16 blocks, +-0"),
    "4k_16x16_pm0": (3840, 2160, 16, 0, 64, "memory-bound small-range case: synthetic 3840x2160 luma, 16x16 blocks, +-0"),
    "1080p_16x16_pm1": (1920, 1080, 16, 1, 256, "memory-bound small-range case: synthetic 1920x1080 luma, 16x16 blocks, +-1"),
    "1080p_16x16_pm2": (1920, 1080, 16, 2, 256, "memory-bound small-range case: synthetic 1920x1080 luma, 16x16 blocks, +-2"),
    "1080p_16x16_pm4": (1920, 1080, 16, 4, 256, "small-range case: synthetic 1920x1080 luma, 16x16 blocks, +-4"),
    "4k_16x16_pm2": (3840, 2160, 16, 2, 64, "memory-bound small-range case: synthetic 3840x2160 luma, 16x16 blocks, +-2"),
    # SSIM-cost full search (src/cpu/main_ssim.c); the first one is that program's default geometry
    "ssim_4k_16x16_pm7": (3840, 2160, 16, 7, 8, "reference SSIM program defaults (main_ssim.c:41-44): synthetic 3840x2160 luma, blk 16, span 7"),
    "ssim_1080p_16x16_pm32": (1920, 1080, 16, 32, 16, "SSIM-cost full search, synthetic 1920x1080 luma, 16x16 blocks, +-32"),
    # SURVEY 8 f-3 / BASELINE configs[3]: fast searches (absent from the reference; parity unpinned)
    "tss_1080p_16x16_pm32": (1920, 1080, 16, 32, 64, "BASELINE configs[3]: three-step search, synthetic 1920x1080 luma, 16x16 blocks, +-32"),
    "diamond_1080p_16x16_pm32": (1920, 1080, 16, 32, 64, "BASELINE configs[3]: diamond search, synthetic 1920x1080 luma, 16x16 blocks, +-32"),
    "diamond_foreman_8x8_pm12": (352, 288, 8, 12, 512, "BASELINE configs[3]: diamond search, Foreman YF2->YF1, 8x8 blocks, +-12"),
}
# (cost, search pattern) of a workload: include/me_b200.h ME_COST_* / ME_SEARCH_*
MODES = {"ssim_4k_16x16_pm7": (1, 0), "ssim_1080p_16x16_pm32": (1, 0), "tss_1080p_16x16_pm32": (0, 1),
         "diamond_1080p_16x16_pm32": (0, 2), "diamond_foreman_8x8_pm12": (0, 2)}
# published reference numbers for the exact same metric, frames/s of the search:
# CPU Beauty 4K 8x8 +-12 = 2350 ms (results/cpu/beauty/2990wx_threadripper_64_cores.txt:12)
PUBLISHED_FPS = {"4k_8x8_pm12": 1000.0 / 2350.0}
METRIC = "1080p_frames_per_sec_full_search_pm32"


def l2_policy(name, pairs):
    """Inputs of one step larger than L2 (126 MB), or an explicit flush between steps."""
    W, H = WORKLOADS[name][0], WORKLOADS[name][1]
    in_bytes = 2 * pairs * H * ((W + 15) & ~15)
    return "inputs larger than L2" if in_bytes >= 160 * 1024 * 1024 else "L2 flushed between steps (192 MB write)"


def workload_config(name, pairs, world):
    """The `config` object of the JSON line -- built by BOTH arms from the workload alone, so the
    driver's same-config check compares equal dicts."""
    from motionestimation_b200.frames import pixel_compares
    W, H, B, R, _, desc = WORKLOADS[name]
    cost, search = MODES.get(name, (0, 0))
    nb = (-(-W // B)) * (-(-H // B))
    return {"workload": name, "desc": desc, "width": W, "height": H, "blk_dim": B, "extra_span": R,
            "pairs_per_gpu_per_step": pairs, "blocks_per_pair": nb,
            "pixel_compares_per_pair": int(pixel_compares(W, H, B, R)),
            "parallelism": f"frame-pair sharding x{world}", "l2": l2_policy(name, pairs),
            "cost": ["mse", "ssim"][cost], "search": ["full", "three_step", "diamond"][search]}


def plan_ingest_helpers(bw, nd, slot_pairs, hp_force=0):
    """Which ranks send part of every submit over a peer GPU's host link (me_b200_set_ingest_helper), and how much.
    bw[r] = copy-only H2D rate of rank r with ALL ranks copying (GB/s), nd[r] = what its kernel consumes.
    Deterministic greedy pairing, the same on every rank: most starved rank first, donor with most spare.  The detour is
    sized to the deficit (rounded, not rounded up) and never takes more than 80 % of the donor's spare link rate: a donor
    driven to 100 % of its own link becomes the slowest rank itself (measured: 4 of 16 pairs over donors with 7.4 GB/s
    spare made the 8-GPU end-to-end rate 4 % worse than no helper at all).  Returns {rank: (helper GPU, pairs per submit)}.
    hp_force > 0 (experiments) fixes the pairs per submit."""
    world = len(bw)
    frac = {r: 1.0 - 0.96 * bw[r] / nd[r] for r in range(world)}      # share of the bytes that must detour
    spare = {r: bw[r] - nd[r] for r in range(world)}
    helpers = {}
    for r in sorted((r for r in range(world) if frac[r] > 0.0), key=lambda r: -frac[r]):
        donors = [d for d in range(world) if d != r and frac[d] <= 0.0 and d not in (h[0] for h in helpers.values())]
        if not donors:
            continue
        d = max(donors, key=lambda d: spare[d])
        hp = max(1, int(round(slot_pairs * frac[r])))
        hp = min(hp, int(0.8 * spare[d] / nd[r] * slot_pairs), slot_pairs - 1)
        if hp_force > 0:
            hp = min(hp_force, slot_pairs - 1)
        if hp >= 1:
            helpers[r] = (d, hp)
    return helpers


def make_batch(name, pairs, rank):
    """Deterministic synthetic batch (pairs, H, W) uint8 x2; a different seed per pair and rank."""
    import motionestimation_b200 as me
    W, H, B, R, _, _ = WORKLOADS[name]
    if "foreman" in name:
        c, r = me.foreman(2), me.foreman(1)
        return np.stack([c] * pairs), np.stack([r] * pairs)
    # a few distinct pairs, repeated: tiled Foreman (real texture/motion) + shifted noise
    base = [me.tiled_frames(W, H, 2, 1), me.tiled_frames(W, H, 4, 1),
            me.shifted_noise_pair(W, H, seed=1234 + rank), me.shifted_noise_pair(W, H, seed=99 + rank, shift=(-11, 7))]
    cur = np.stack([base[i % len(base)][0] for i in range(pairs)])
    ref = np.stack([base[i % len(base)][1] for i in range(pairs)])
    return cur, ref


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md recipe),
    read through NVML every few ms (nvidia-smi itself is too slow for sub-second regions)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.max_mhz, self.err = index, [], False, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES when mapping the torch index to an NVML index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while True:   # at least one sample even when the timed region is shorter than the start-up
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((float(mhz), int(reasons), watts))
                if self.stop_flag:
                    break
                time.sleep(0.004)
        except Exception as ex:  # pragma: no cover
            self.err = repr(ex)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: %s" % self.err]}
        sm = sorted(r[0] for r in self.rows)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}
        seen = 0
        for r in self.rows:
            seen |= r[1]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz,
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows),
                "reasons": [n for n, b in bits.items() if seen & b]}


def load_ref(mode=(0, 0)):
    """oracle/_ref (the unmodified reference, prebuilt) -- test/baseline infrastructure."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Ref, RefSsim, Oracle
    if mode[1]:
        return Oracle(), "port"          # fast searches do not exist in the reference
    if mode[0] == 1:
        return (RefSsim(), "reference") if RefSsim.available() else (Oracle(), "port")
    if Ref.available():
        return Ref(), "reference"
    return Oracle(), "port"


def cpu_reference_rate(name, budget_s, steps=1, warmup=0):
    """Time the reference CPU search (its own thread-pool dispatch, main.c:144-158, 100 threads)
    on a bounded sample of the workload: whole block rows from the middle of the first pair.
    Returns (frames/s, blocks/s, dict describing the sample)."""
    W, H, B, R, _, _ = WORKLOADS[name]
    cur, ref = make_batch(name, 1, 0)
    cur, ref = cur[0], ref[0]
    mode = MODES.get(name, (0, 0))
    lib, kind = load_ref(mode)
    nbx, nby = -(-W // B), -(-H // B)
    nb = nbx * nby

    def run(b0, b1):
        if kind == "reference" and mode == (0, 0):
            sec, _ = lib.search_pool(cur, ref, B, R, b0, b1, pool_threads=100)
            return sec
        t0 = time.perf_counter()
        if mode[1]:
            lib.search_fast(cur, ref, B, R, mode[1], b0, b1, nthreads=os.cpu_count())
        elif mode[0] == 1 and kind == "reference":
            lib.search(cur, ref, B, R, b0, b1, nthreads=os.cpu_count())   # unmodified findBestBlkSSIM
        elif mode[0] == 1:
            lib.search_ssim(cur, ref, B, R, b0, b1, nthreads=os.cpu_count())
        else:
            lib.search(cur, ref, B, R, b0, b1, nthreads=os.cpu_count())
        return time.perf_counter() - t0

    # probe with two interior block rows, then size the sample to the time budget
    mid = (nby // 2) * nbx
    t_probe = run(mid, min(nb, mid + 2 * nbx))
    per_row = t_probe / 2
    rows = int(max(1, min(nby, budget_s / max(per_row, 1e-6) / max(1, steps + warmup))))
    if rows >= nby:
        b0, b1, rows = 0, nb, nby
    else:
        r0 = max(0, nby // 2 - rows // 2)
        b0, b1 = r0 * nbx, (r0 + rows) * nbx
    for _ in range(warmup):
        run(b0, b1)
    ts = [run(b0, b1) for _ in range(steps)]
    t = float(np.mean(ts))
    blocks_s = (b1 - b0) / t
    cores = os.cpu_count() or 1
    if mode == (0, 0):
        how = ("reference thread pool of 100 (main.c:144), built -O2 from the unmodified sources "
               "(src/cpu/run.sh:4 uses no -O)")
    elif mode[1]:
        how = ("CPU definition of the fast search (no reference implementation exists), one pthread per core")
    else:
        how = ("unmodified findBestBlkSSIM (main_ssim.c:16, ssim.c) built -O2, blocks split over one pthread per "
               "core (the reference program itself runs them on one thread, main_ssim.c:67-77)")
    return blocks_s / nb, blocks_s, {
        "kind": kind, "cores": cores, "threads": 100 if (kind == "reference" and mode == (0, 0)) else cores,
        "sample": f"{rows} of {nby} block rows ({b1 - b0} blocks) of one {W}x{H} pair, B={B} R={R}, "
                  f"{steps} timed run(s) of {t * 1e3:.0f} ms; {how}",
        "ms_per_step": t * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    W, H, B, R, pairs, desc = WORKLOADS[name]
    if args.pairs > 0:
        pairs = args.pairs
    fps, blocks_s, info = cpu_reference_rate(name, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC if name == "1080p_16x16_pm32" else name + "_frames_per_sec",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(name, pairs, args.gpus),
        "blocks_per_s": blocks_s,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_check(name, cur_np, ref_np, B, R, mode, dev_out):
    """Every block (MV, integer SSD / found flag, float score bits) of one pair of the timed batch --
    the shifted-noise pair when the batch has one -- against the oracle on all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle
    orc = Oracle()
    p = 2 if cur_np.shape[0] > 2 else 0
    t0 = time.perf_counter()
    if mode[1]:
        exp, _ = orc.search_fast(cur_np[p], ref_np[p], B, R, mode[1], nthreads=os.cpu_count())
    elif mode[0] == 1:
        exp = orc.search_ssim(cur_np[p], ref_np[p], B, R, nthreads=os.cpu_count())
    else:
        exp = orc.search(cur_np[p], ref_np[p], B, R, nthreads=os.cpu_count())
    secs = time.perf_counter() - t0
    got = {k: v[p].cpu().numpy() for k, v in dev_out.items()}
    ok = (np.array_equal(got["mvx"], exp["mvx"]) and np.array_equal(got["mvy"], exp["mvy"]) and
          np.array_equal(got["ssd"].view(np.uint32), exp["ssd"]) and
          np.array_equal(got["score"].view(np.uint32), exp["score"].view(np.uint32)))
    if not ok:
        raise SystemExit("bench.py: the timed batch does NOT match the oracle (pair %d of %s)" % (p, name))
    return {"checked": True, "pair": p, "blocks": int(exp.shape[0]),
            "fields": ["mvx", "mvy", "ssd", "score bits"], "oracle_seconds": secs,
            "checker": "oracle/me_oracle*.c (CPU restatement of main.c:18-82, pinned against the unmodified "
                       "reference), all host cores, outside the timed region"}


def band_split_leg(me, torch, dist, local, rank, world):
    """ONE 4K 16x16 pair split by cost-balanced block-row bands over the ranks.  Two ways to complete
    the field: one NCCL all_gather of the packed arrays, or none -- the search kernel stores every
    block into all ranks' peer-mapped copies (NVLink) and a device-side flag barrier follows.
    Both are checked on every rank against the unsharded search of that rank."""
    from motionestimation_b200 import sharding
    W, H, B = 3840, 2160, 16
    out = {"workload": "one 3840x2160 pair, 16x16 blocks, split by block-row bands over %d GPUs" % world,
           "unit": "ms", "timing": "CUDA events, median of 7, max over ranks"}
    pairs = [me.tiled_frames(W, H), me.shifted_noise_pair(W, H, seed=7)]
    d = [(torch.from_numpy(c).cuda(), torch.from_numpy(r).cuda()) for c, r in pairs]

    def timed(fn):
        ts = []
        for i in range(7):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn(i)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), res

    def same(res, full):
        return all(bool(torch.equal(res[k].view(torch.int32), full[k].view(torch.int32))) for k in ("mvx", "mvy", "ssd", "score"))

    for R in (32, 64):
        with me.Estimator(W, H, B, R, device=local) as est:
            nb = est.num_blocks
            full = []
            for c, r in d:
                o = {k: torch.zeros((1, nb), dtype=torch.int32, device="cuda") for k in ("mvx", "mvy", "ssd")}
                o["score"] = torch.zeros((1, nb), dtype=torch.float32, device="cuda")
                est.search_device(c, r, W, W * H, 1, o["mvx"], o["mvy"], o["ssd"], o["score"])
                full.append(o)
            torch.cuda.synchronize()
            t1, _ = timed(lambda i: est.search_device(d[0][0], d[0][1], W, W * H, 1, full[0]["mvx"], full[0]["mvy"],
                                                      full[0]["ssd"], full[0]["score"]))
            # the inputs alternate between two different pairs, so a stale or half-written field shows
            for i in range(3):
                sharding.search_banded(est, d[i & 1][0], d[i & 1][1], W, W * H, 1)
            ok_n = True
            tn, _ = timed(lambda i: sharding.search_banded(est, d[i & 1][0], d[i & 1][1], W, W * H, 1))
            for i in range(4):
                ok_n &= same(sharding.search_banded(est, d[i & 1][0], d[i & 1][1], W, W * H, 1), full[i & 1])
            field = sharding.PeerField(est, 1)
            for i in range(4):
                sharding.search_banded_peer(est, field, d[i & 1][0], d[i & 1][1], W, W * H, 1, check=False)
            field.check()
            tp, _ = timed(lambda i: sharding.search_banded_peer(est, field, d[i & 1][0], d[i & 1][1], W, W * H, 1,
                                                                check=False))
            ok_p = True
            for i in range(6):   # back to back, no host synchronisation in between: exercises the double buffering
                res = sharding.search_banded_peer(est, field, d[i & 1][0], d[i & 1][1], W, W * H, 1, check=False)
                ok_p &= same(res, full[i & 1])
            field.check()
            torch.cuda.synchronize()
            field.close()
            flag = torch.tensor([int(ok_n), int(ok_p)], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            out["pm%d" % R] = {"unsharded_ms": t1, "nccl_all_gather_ms": tn, "peer_stores_ms": tp,
                               "identical_on_all_ranks_and_equal_to_unsharded": bool(flag[0].item() and flag[1].item())}
            if not (flag[0].item() and flag[1].item()):
                raise SystemExit("bench.py: band-sharded field differs from the unsharded search")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p_16x16_pm32", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-band-split", action="store_true")
    ap.add_argument("--no-post", action="store_true")
    ap.add_argument("--sustained-s", type=float, default=3.0, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--dropin-calls", type=int, default=50, help="me_b200_search calls of the e2e_dropin leg (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import motionestimation_b200 as me

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the search has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["ME_B200_DEVICE"] = str(local)     # device of the implicit context of me_b200_search
    # one slice of the host cores per rank (the feeding threads of eight ranks otherwise migrate
    # over the whole socket); BENCH_AFFINITY=0 switches it off
    all_cpus = sorted(os.sched_getaffinity(0))
    affinity = None
    if world > 1 and os.environ.get("BENCH_AFFINITY", "1") != "0" and len(all_cpus) >= 2 * world:
        per = len(all_cpus) // world
        affinity = all_cpus[local * per:(local + 1) * per]
        os.sched_setaffinity(0, affinity)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)

    name = args.workload
    W, H, B, R, pairs, desc = WORKLOADS[name]
    if args.pairs > 0:
        pairs = args.pairs
    pitch = (W + 15) & ~15
    cur_np, ref_np = make_batch(name, pairs, rank)

    slot_pairs = max(1, pairs // me.lib.ME_B200_MAX_SLOTS)
    cost, search = MODES.get(name, (0, 0))
    est = me.Estimator(W, H, B, R, device=local, max_pairs=slot_pairs, cost=cost, search=search)
    nb = est.num_blocks
    pc_pair = est.pixel_compares

    # ---- device-resident arm -------------------------------------------------------------
    d_cur = torch.zeros((pairs, H, pitch), dtype=torch.uint8, device="cuda")
    d_ref = torch.zeros_like(d_cur)
    d_cur[:, :, :W] = torch.from_numpy(cur_np).cuda()
    d_ref[:, :, :W] = torch.from_numpy(ref_np).cuda()
    d_mvx = torch.zeros((pairs, nb), dtype=torch.int32, device="cuda")
    d_mvy = torch.zeros_like(d_mvx)
    d_ssd = torch.zeros_like(d_mvx)
    d_score = torch.zeros((pairs, nb), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()
    in_bytes = 2 * pairs * H * pitch
    # L2 hygiene: either the step's inputs exceed L2 (126 MB) or we flush it between steps
    flush = None
    if in_bytes < 160 * 1024 * 1024:
        flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step():
        est.search_device(d_cur, d_ref, pitch, H * pitch, pairs, d_mvx, d_mvy, d_ssd, d_score,
                          stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = est.launch_count
    evs = []
    sync_all()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2.0)
    launches = est.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = float(sum(step_ms))                      # device time of the K steps on this rank
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    fps = world * pairs * args.steps / (dev_ms_max * 1e-3)

    mvx0 = d_mvx[0].cpu().numpy()   # compared with the end-to-end arm below

    # ---- end-to-end arm: host buffers through the C ABI, copies inside the timed region ---
    lib = me.load_library()
    n = W * H
    host_allocs = []

    def pinned(shape, dtype, upload_only=False):
        """numpy view of memory from me_b200_host_alloc_ex (cudaHostAlloc, allocated by this rank's own
        thread after its affinity was set -> first touched on the rank's cores).  BENCH_WC=1: the
        upload-only frame buffers are write-combined."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        wc = upload_only and os.environ.get("BENCH_WC", "0") == "1"
        ptr = lib.me_b200_host_alloc_ex(nbytes, me.ME_HOST_WRITE_COMBINED if wc else 0)
        if not ptr:
            raise SystemExit("bench.py: me_b200_host_alloc_ex failed")
        host_allocs.append(ptr)
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr)).view(dtype).reshape(shape)

    h_cur = pinned((pairs, n), np.uint8, True)
    h_ref = pinned((pairs, n), np.uint8, True)
    h_cur[:] = cur_np.reshape(pairs, n)
    h_ref[:] = ref_np.reshape(pairs, n)
    h_mvx = pinned((pairs, nb), np.int32)
    h_mvy = pinned((pairs, nb), np.int32)
    h_ssd = pinned((pairs, nb), np.int32)
    h_mvx[:] = 0
    fsz, osz = n, nb * 4    # bytes per pair of a frame array / an output array
    p_cur, p_ref, p_mvx, p_mvy, p_ssd = (a.ctypes.data for a in (h_cur, h_ref, h_mvx, h_mvy, h_ssd))
    nslots = min(me.lib.ME_B200_MAX_SLOTS, pairs // slot_pairs)

    def e2e_step():
        inflight = [False] * nslots
        done, k = 0, 0
        while done < pairs:
            s = k % nslots
            if inflight[s]:
                est.wait(s)
            npp = min(slot_pairs, pairs - done)
            est.submit_ptr(s, p_cur + done * fsz, p_ref + done * fsz, npp, p_mvx + done * osz,
                           p_mvy + done * osz, p_ssd + done * osz, 0)
            inflight[s] = True
            done += npp
            k += 1
        for s in range(nslots):
            if inflight[s]:
                est.wait(s)

    def e2e_stream(nsteps):
        """K steps streamed through the slot ring the way the API is meant to be used: the upload of
        the next batch (also the first batch of the next step) overlaps the search of the current
        one; every batch's H2D, search and D2H lie inside the timed region."""
        inflight = [False] * nslots
        k = 0
        for _ in range(nsteps):
            done = 0
            while done < pairs:
                s = k % nslots
                if inflight[s]:
                    est.wait(s)
                npp = min(slot_pairs, pairs - done)
                est.submit_ptr(s, p_cur + done * fsz, p_ref + done * fsz, npp, p_mvx + done * osz,
                               p_mvy + done * osz, p_ssd + done * osz, 0)
                inflight[s] = True
                done += npp
                k += 1
        for s in range(nslots):
            if inflight[s]:
                est.wait(s)

    # ---- ingest routing (N > 1): not every GPU of a box has a full-speed host link (profiles/h2d_probe_r02.txt:
    # four GPUs of this pool's 8-GPU boxes share one PCIe uplink).  Every rank measures its copy-only H2D rate
    # with ALL ranks copying at once; a rank whose link cannot feed its kernel sends the last pairs of every
    # submit over a peer GPU with spare link capacity and NVLink (me_b200_set_ingest_helper).
    ingest = None
    if world > 1 and os.environ.get("BENCH_INGEST_HELPER", "1") != "0":
        pb = 64 << 20
        ph = pinned((pb,), np.uint8, True)
        ph[:] = 1
        pt = torch.from_numpy(ph)
        pd = torch.empty(pb, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            pd.copy_(pt, non_blocking=True)
        sync_all()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(16):
            pd.copy_(pt, non_blocking=True)
        q1.record()
        torch.cuda.synchronize()
        my_bw = 16 * pb / (q0.elapsed_time(q1) * 1e-3) / 1e9
        need = 2 * n * (pairs * args.steps / (dev_ms * 1e-3)) / 1e9        # GB/s this rank's kernel consumes
        t = torch.tensor([my_bw, need], dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        bw = [float(v[0].item()) for v in allv]
        nd = [float(v[1].item()) for v in allv]
        del pd
        helpers = plan_ingest_helpers(bw, nd, slot_pairs, int(os.environ.get("BENCH_INGEST_HP", "0")))
        helper_error = None
        if rank in helpers:
            try:   # (local rank == device index on one node; needs every GPU visible to every rank + peer access)
                if torch.cuda.device_count() <= helpers[rank][0]:
                    raise RuntimeError("GPU %d is not visible to rank %d" % (helpers[rank][0], rank))
                est.set_ingest_helper(helpers[rank][0], helpers[rank][1])
            except Exception as ex:   # the direct path still works, only slower
                helper_error = repr(ex)
        ingest = {"h2d_gbs_all_ranks_copying": [round(b, 1) for b in bw], "kernel_needs_gbs": round(nd[0], 1),
                  "helpers": {str(r): {"via_gpu": h[0], "pairs_per_submit": h[1], "of": slot_pairs}
                              for r, h in helpers.items()},
                  "rank0_helper_error": helper_error,
                  "api": "me_b200_set_ingest_helper: host -> helper GPU (its PCIe link) -> NVLink peer copy"}
        sync_all()

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    sync_all()
    t0 = time.perf_counter()
    e2e_stream(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = world * pairs * args.steps / float(t[0].item())
    e2e_sync_fps = world * pairs * args.steps / float(t[1].item())
    assert np.array_equal(h_mvx[0], mvx0), "e2e and device-resident paths disagree"

    # ---- same, for a video sequence: consecutive pairs share frames, each frame crosses PCIe once
    seq_fps = None
    if slot_pairs >= 1:
        seq = np.concatenate([ref_np[:1], cur_np[:1]] * ((slot_pairs + 2) // 2))[:slot_pairs + 1]
        h_seq = pinned((slot_pairs + 1, n), np.uint8, True)
        h_seq[:] = np.ascontiguousarray(seq).reshape(slot_pairs + 1, n)

        def seq_step():
            for s_ in range(nslots):
                est.submit_sequence_ptr(s_, h_seq.ctypes.data, slot_pairs + 1, p_mvx + s_ * slot_pairs * osz,
                                        p_mvy + s_ * slot_pairs * osz, p_ssd + s_ * slot_pairs * osz, 0)
            for s_ in range(nslots):
                est.wait(s_)

        for _ in range(2):
            seq_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            seq_step()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seq_fps = world * nslots * slot_pairs * args.steps / float(t.item())

    # ---- sustained leg: back-to-back steps for >= 3 s (no flush, no events in between) with the
    # clock sampler running -- shows whether the clocks of the short timed region hold
    sustained = None
    if args.sustained_s > 0:
        per_step_s = max(dev_ms / args.steps * 1e-3, 1e-5)
        nsus = int(max(args.steps, min(200000, args.sustained_s / per_step_s + 1)))
        sus_sampler = ClockSampler(local)
        sync_all()
        if rank == 0:
            sus_sampler.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for i in range(nsus):
            step()
        s1.record(stream)
        sync_all()
        sus_sampler.stop_flag = True
        if rank == 0:
            sus_sampler.join(timeout=2.0)
        t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t.item())
        if rank == 0:
            sustained = {"value": world * pairs * nsus / (sus_ms * 1e-3), "unit": "frames/s", "steps": nsus,
                         "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / nsus,
                         "l2": "no flush between steps (inputs %s L2)" % ("exceed" if flush is None else "fit"),
                         "clocks": sus_sampler.summary()}

    # ---- the literal reference seam: me_b200_search(predictionFrame*, const int*, int) on `int`
    # frames (main.c:132-158), one blocking call per frame pair
    dropin = None
    if cost == 0 and search == 0 and args.dropin_calls > 0:
        ci = np.ascontiguousarray(cur_np[2 % pairs].astype(np.int32).ravel())
        ri = np.ascontiguousarray(ref_np[2 % pairs].astype(np.int32).ravel())
        pf = me.create_prediction_frame(ci, W, H, B)
        for _ in range(3):
            me.search_prediction_frame(pf, ri, R)
        ncalls = args.dropin_calls
        sync_all()
        t0 = time.perf_counter()
        for _ in range(ncalls):
            me.search_prediction_frame(pf, ri, R)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mv = np.array([(pf.blks[i].motion_vectorX, pf.blks[i].motion_vectorY) for i in range(0, nb, max(1, nb // 257))])
        ref_mv = np.stack([d_mvx[2 % pairs].cpu().numpy(), d_mvy[2 % pairs].cpu().numpy()], 1)[::max(1, nb // 257)]
        assert np.array_equal(mv, ref_mv), "drop-in and device-resident paths disagree"
        dropin = {"value": world * ncalls / float(t.item()), "unit": "frames/s", "ms_per_call": float(t.item()) / ncalls * 1e3,
                  "calls": ncalls, "h2d_bytes_per_call": 2 * n, "host_int_bytes_per_call": 8 * n,
                  "api": "me_b200_search(predictionFrame*, const int*, int): blocking, int frames narrowed to u8 by "
                         "the library's worker threads chunk by chunk while the chunks upload, search, MVs "
                         "written into the reference's block structs"}

    # ---- the stage right after the path (main.c:160-171 / utils.c:94-164), batched on the
    # device: 5 output planes + the PSNR integers of every pair of the step's batch.  HBM-bound.
    post = None
    if cost == 0 and search == 0 and not args.no_post:
        pp = min(pairs, max(1, (2 << 30) // (5 * W * H)))          # at most 2 GB of output planes
        out5 = torch.empty((pp, 5 * H * W), dtype=torch.uint8, device="cuda")
        sq = torch.zeros(pp, dtype=torch.int64, device="cuda")
        mxv = torch.zeros(pp, dtype=torch.int32, device="cuda")
        post_bytes = pp * (2 + 5) * W * H

        def post_step():
            est.postprocess_device_batch(d_cur, d_ref, pitch, H * pitch, pp, d_mvx, d_mvy, out5, 5 * W * H, sq, mxv,
                                         stream.cuda_stream)
        for _ in range(3):
            post_step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            if flush is not None and post_bytes < 160 * 1024 * 1024:
                flush.fill_(1)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            post_step()
            p1.record(stream)
            torch.cuda.synchronize()
            ts.append(p0.elapsed_time(p1))
        pms = float(np.median(ts))
        post = {"value": pp / (pms * 1e-3), "unit": "frames/s", "pairs": pp, "ms": pms,
                "algorithmic_bytes": post_bytes, "achieved_gbs": post_bytes / (pms * 1e-3) / 1e9,
                "api": "me_b200_postprocess_device_batch: ref, cur, motion-compensated, |ref-cur|, |mc-cur| planes "
                       "+ sum (mc-cur)^2 and max pixel per pair; 2 B read + 5 B written per pixel"}
        del out5

    # ---- one very large frame split by block-row bands over the ranks: NCCL gather vs
    # peer-mapped fields; every rank checks its complete field against the unsharded search
    band_split = None
    if world > 1 and not args.no_band_split:
        band_split = band_split_leg(me, torch, dist, local, rank, world)

    # ---- parity of the timed batch: every block of one pair against the oracle (outside the timed
    # region; the oracle is only the checker)
    parity = None
    if rank == 0 and not args.no_parity_check:
        parity = parity_check(name, cur_np, ref_np, B, R, (cost, search),
                              {"mvx": d_mvx, "mvy": d_mvy, "ssd": d_ssd, "score": d_score})

    if rank == 0:
        # ---- roofline of the search kernel (integer pipes) -----------------
        pair_rate, mhz = 0.0, 0.0
        # MSE: VABSDIFF4 + IDP.4A pairs (2 ops per 4 pixels); SSIM: the only per-pixel work of a
        # candidate is the dot product, 1 IDP.4A per 4 pixels, against the IDP.4A rate alone
        which, per_pc, what = (0, 0.25, "IDP.4A") if cost == 1 else (2, 0.5, "VABSDIFF4+IDP.4A pairs")
        for _ in range(3):
            r_, m_ = me.int_peak(which, iters=4000, device=local)
            if r_ > pair_rate:
                pair_rate, mhz = r_, m_
        step_s = dev_ms / args.steps * 1e-3
        fast_evals = None
        if search:
            # a fast search evaluates a data-dependent handful of candidates per block
            ev0 = est.fast_evaluations
            step()
            fast_evals = est.fast_evaluations - ev0
            pc_step = fast_evals * B * B   # upper bound (partial edge blocks are smaller)
        else:
            pc_step = pc_pair * pairs
        lane_instr = per_pc * pc_step                            # algorithmic lane-instructions per step
        achieved = lane_instr / step_s
        alg_bytes = pairs * (2 * W * H + 16 * nb)                # u8 cur + ref read once, 16 B/block out
        clocks = sampler.summary()
        traffic = None   # DRAM bytes of the search kernel per launch, from the committed ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)[name]["bytes_per_pair"] * pairs
        except Exception:
            pass
        # measured copy bandwidth of this pool's B200s (driver-written), else the profiling guide's fallback
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
        roofline = {"bound": "int_alu", "achieved": achieved / 1e12, "peak": pair_rate / 1e12,
                    "unit": "T lane-instr/s", "frac": achieved / pair_rate if pair_rate else None,
                    "traffic": traffic,
                    "peak_source": "measured live: me_b200_int_peak(%s) at %.0f MHz" % (what, mhz),
                    "frac_of_single_pipe_peak": (achieved / (pair_rate / 2) if pair_rate else None) if cost == 0 else None,
                    "note": ("fast search: a few dependent steps per block, latency bound by design; "
                             "the fraction only says how little arithmetic a fast search needs") if search else None,
                    "hbm": {"achieved_gbs": alg_bytes / step_s / 1e9, "peak_gbs": hbm_peak,
                            "algorithmic_bytes_per_step": alg_bytes}}
        if R <= 2 and not cost and not search:
            # a handful of candidates per block (+-0..+-2 are the bandwidth-bound
            # ranges with u8 frames): a streaming pass, bound by HBM
            roofline = {"bound": "hbm", "achieved": alg_bytes / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / step_s / 1e9 / hbm_peak, "traffic": traffic,
                        "peak_source": hbm_src,
                        "int_alu": {"achieved_t_lane_instr_s": achieved / 1e12, "peak": pair_rate / 1e12}}
        line = {
            "metric": METRIC if name == "1080p_16x16_pm32" else name + "_frames_per_sec",
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (fps / PUBLISHED_FPS[name]) if name in PUBLISHED_FPS else None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(name, pairs, world),
            "kernel": ("warp-per-block fast search" if search else "ssim tiled + statistics pre-pass" if cost
                       else {1: "generic", 2: "tiled", 3: "direct"}[est.kernel_in_use]),
            "fallback_launches": est.fallback_launches,
            "parity_checked": bool(parity and parity["checked"]), "parity": parity,
            "blocks_per_s": fps * nb, "pixel_compares_per_s": pc_step / step_s,
            "candidate_evaluations_per_s": (fast_evals / step_s) if fast_evals is not None else None,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 2 * pairs * n,
                    "d2h_bytes_per_step": 3 * pairs * nb * 4,
                    "api": "me_b200_submit/me_b200_wait, pinned u8 host frames, %d slots x %d pairs, the K steps "
                           "streamed through the slot ring (upload of batch i+1 overlaps the search of batch i)"
                           % (nslots, slot_pairs),
                    "step_synchronous": e2e_sync_fps,
                    "ingest_routing": ingest,
                    # host->device traffic the streamed run actually sustained (per GPU)
                    "h2d_gbs_per_gpu": 2 * pairs * n * (e2e_fps / world / pairs) / 1e9},
            "e2e_sequence": {"value": seq_fps, "unit": "frames/s",
                             "h2d_bytes_per_step": nslots * (slot_pairs + 1) * n,
                             "api": "me_b200_submit_sequence: pair i = frame i+1 vs frame i, each frame uploaded once"},
            "e2e_dropin": dropin,
            "post": (dict(post, peak_gbs=hbm_peak, frac=post["achieved_gbs"] / hbm_peak, peak_source=hbm_src)
                     if post else None),
            "sustained": sustained,
            "band_split": band_split,
            "host": {"cpus": len(all_cpus), "affinity_of_rank0": affinity,
                     "pinned": "me_b200_host_alloc_ex" + (" (write-combined frames)" if os.environ.get("BENCH_WC") == "1" else "")},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if est.fallback_launches:
            raise SystemExit("bench.py: %d launches fell back to the generic kernel: %s"
                             % (est.fallback_launches, lib.me_b200_last_error(est._h).decode()))
        if not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)   # the CPU baseline uses every host core
            cfps, cblocks, info = cpu_reference_rate(name, budget_s=30.0, steps=3, warmup=1)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": info["cores"],
                                    "kind": info["kind"], "sample": info["sample"], "blocks_per_s": cblocks}
        print(json.dumps(line), flush=True)
    est.close()
    lib.me_b200_release_cached()
    for ptr in host_allocs:
        lib.me_b200_host_free(ptr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
