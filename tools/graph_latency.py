"""Single-pair search latency on device-resident frames: plain launches vs the same call captured
once into a CUDA graph and replayed (the call is stream-ordered end to end -- stream-ordered scratch,
tensor maps passed by value -- so it can be captured).
usage: python tools/graph_latency.py [W H B R]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    W, H, B, R = (map(int, sys.argv[1:5]) if len(sys.argv) > 4 else (352, 288, 8, 12))
    cur8, ref8 = (me.foreman(2), me.foreman(1)) if (W, H) == (352, 288) else me.tiled_frames(W, H)
    pitch = (W + 15) & ~15
    cur = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda")
    ref = torch.zeros_like(cur)
    cur[:, :W] = torch.from_numpy(cur8).cuda()
    ref[:, :W] = torch.from_numpy(ref8).cuda()
    with me.Estimator(W, H, B, R) as est:
        nb = est.num_blocks
        out = [torch.zeros((nb,), dtype=torch.int32, device="cuda") for _ in range(4)]
        exp = est.search_u8(cur8, ref8)
        s = torch.cuda.Stream()

        def call():
            est.search_device(cur, ref, pitch, pitch * H, 1, out[0], out[1], out[2], out[3].view(torch.float32),
                              s.cuda_stream)

        def timeit(fn, reps=200):
            with torch.cuda.stream(s):
                for _ in range(10):
                    fn()
                s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(reps):
                    fn()
                e1.record(s)
                s.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3

        t_plain = timeit(call)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        for o in out:
            o.zero_()
        t_graph = timeit(g.replay)
        s.synchronize()
        ok = (np.array_equal(out[0].cpu().numpy(), exp["mvx"][0]) and np.array_equal(out[1].cpu().numpy(), exp["mvy"][0])
              and np.array_equal(out[2].cpu().numpy().view(np.uint32), exp["ssd"][0]))
        print(f"{W}x{H} B={B} R={R}: {t_plain:.1f} us per search with plain launches, {t_graph:.1f} us replaying a "
              f"captured CUDA graph; graph result identical: {ok}")
        assert ok


if __name__ == "__main__":
    main()
