"""GPU tests of the peer-field band path (me_b200_search_device_band_peers / me_b200_peer_barrier)
on ONE GPU: the "peer" fields are further allocations on the same device, which exercises the same
stores the multi-GPU run does over NVLink (tools/band_check.py runs the real thing on 2/4/8 GPUs)."""
import ctypes as C

import numpy as np
import pytest

import motionestimation_b200 as me
from oracle_binding import Oracle

pytestmark = pytest.mark.gpu


def _field(est, base, entries):
    a = 4 * entries
    return me.Field(base, base + a, base + 2 * a, base + 3 * a)


def _read(est, base, entries, nb, npairs):
    import torch
    from motionestimation_b200.sharding import _RawCuda
    t = torch.as_tensor(_RawCuda(base, (4, npairs, nb), "<i4"), device="cuda").cpu().numpy()
    return {"mvx": t[0], "mvy": t[1], "ssd": t[2].view(np.uint32), "score": t[3].view(np.float32)}


@pytest.mark.parametrize("B,R,W,H,kw", [
    (16, 32, 352, 288, {}),                      # tuned kernel: stores into the peers from the publish step
    (16, 12, 200, 100, {}),                      # tuned rows + a 4-pixel bottom row on the generic kernel
    (8, 2, 128, 72, {}),                         # small-span kernel: peer store kernel
    (5, 3, 41, 23, {}),                          # generic kernel
    (16, 7, 96, 80, {"cost": me.ME_COST_SSIM}),  # SSIM cost
    (8, 12, 100, 60, {"search": me.ME_SEARCH_DIAMOND}),
])
def test_band_peers_on_one_gpu(B, R, W, H, kw):
    import torch
    orc = Oracle()
    pairs = [me.shifted_noise_pair(W, H, seed=3, shift=(2, -1), cell=4), me.random_pair(W, H, 4)]
    cur = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    ref = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    with me.Estimator(W, H, B, R, max_pairs=2, **kw) as est:
        nb, nby = est.num_blocks, est.blocks_y
        entries = 2 * nb
        bases = [est.device_alloc(16 * entries + 256) for _ in range(3)]
        assert len(est.ipc_export(bases[0])) == me.lib.ME_B200_IPC_HANDLE_BYTES
        local, peers = _field(est, bases[0], entries), [_field(est, b, entries) for b in bases[1:]]
        # two bands, as two ranks would issue them (here both from one process)
        mid = nby // 2
        est.search_device_band_peers(cur, ref, W, W * H, 2, 0, mid, local, peers)
        est.search_device_band_peers(cur, ref, W, W * H, 2, mid, nby, local, peers)
        # a one-rank barrier: raises and observes its own flag
        flags = [bases[0] + 16 * entries]
        est.peer_barrier(flags, 0, 1)
        est.peer_barrier(flags, 0, 2)
        assert not est.peer_barrier_timed_out()
        got = [_read(est, b, entries, nb, 2) for b in bases]
        for b in bases:
            est.device_free(b)
    for p, (c, r) in enumerate(pairs):
        if kw.get("cost"):
            exp = orc.search_ssim(c, r, B, R)
        elif kw.get("search"):
            exp, _ = orc.search_fast(c, r, B, R, kw["search"])
        else:
            exp = orc.search(c, r, B, R)
        for g in got:
            assert np.array_equal(g["mvx"][p], exp["mvx"]) and np.array_equal(g["mvy"][p], exp["mvy"])
            assert np.array_equal(g["ssd"][p], exp["ssd"])
            assert np.array_equal(g["score"][p].view(np.uint32), exp["score"].view(np.uint32))


def test_peer_barrier_times_out_instead_of_hanging():
    """A flag nobody raises: the barrier kernel gives up after the time-out and reports it."""
    with me.Estimator(64, 48, 8, 4) as est:
        a, b = est.device_alloc(256), est.device_alloc(256)
        est.peer_barrier([a, b], 0, 1, timeout_ms=20)   # 'rank 1' never signals
        assert est.peer_barrier_timed_out()
        assert not est.peer_barrier_timed_out()        # the status is cleared by the query
        est.device_free(a)
        est.device_free(b)


def test_peer_argument_validation():
    lib = me.load_library()
    with me.Estimator(64, 48, 8, 4) as est:
        f = me.Field(None, None, None, None)
        assert lib.me_b200_search_device_band_peers(est._h, None, None, 64, 0, 1, 0, 1, C.byref(f), None, 0,
                                                    None) == me.ME_ERR_INVALID_ARG
        assert lib.me_b200_peer_barrier(est._h, None, 2, 0, 1, 10, None) == me.ME_ERR_INVALID_ARG
        arr = (C.c_void_p * 2)(None, None)
        assert lib.me_b200_peer_barrier(est._h, arr, 2, 0, 1, 10, None) == me.ME_ERR_INVALID_ARG
        assert lib.me_b200_peer_barrier(est._h, arr, 9, 0, 1, 10, None) == me.ME_ERR_INVALID_ARG
