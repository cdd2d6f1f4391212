"""Multi-GPU partitioning of the search.

The reference has no multi-device code at all (its only parallelism is one
thread-pool job per block, ``src/cpu/main.c:144-158``); blocks -- and therefore
frame pairs -- are independent, which gives two shardings, one process per GPU:

* **by frame pair** (the batched workloads): rank r owns a contiguous slice of the
  batch and writes its own slice of the motion field.  No data-path collective.
* **by block-row band** (one very large frame): rank r searches block rows
  ``[band_rows(...)]`` of every pair with ``me_b200_search_device_band`` -- the kernel
  reads the reference rows of the band +- R straight from the rank's copy of the
  frame, so there is no halo exchange -- and ONE collective at the end gathers the
  per-band slices of the field (``all_gather`` over NCCL on GPUs; gloo in CPU tests).

  :class:`PeerField` + :func:`search_banded_peer` do the same WITHOUT a collective: every rank
  maps the field buffers of its peers (CUDA IPC over NVLink/NVSwitch) and the search kernel
  stores each finished block into all copies; a device-side flag barrier replaces the gather.

torch.distributed is plumbing here: rendezvous, the exchange of the IPC handles and the gather.
The search is always the CUDA library.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def pair_slice(npairs: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a batch of frame pairs owned by `rank`
    (sizes differ by at most one, earlier ranks take the remainder)."""
    if world < 1 or not 0 <= rank < world or npairs < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(npairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def band_rows(nby: int, world: int, rank: int, weights: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """Block rows [begin, end) of rank `rank` when one frame is split into bands.
    With `weights` (one cost per block row, e.g. from :func:`row_costs`) the cut
    points balance cost instead of row count (border rows are cheaper: their
    windows are clamped, main.c:73-76)."""
    if world < 1 or not 0 <= rank < world or nby < 0:
        raise ValueError("bad partition arguments")
    if weights is None:
        return pair_slice(nby, world, rank)
    if len(weights) != nby:
        raise ValueError("one weight per block row expected")
    total = float(sum(weights))
    cuts = [0]
    acc, row = 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while row < nby and acc + weights[row] / 2.0 <= target:
            acc += weights[row]
            row += 1
        cuts.append(row)
    cuts.append(nby)
    for i in range(1, len(cuts)):           # monotone, never empty unless nby < world
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def row_costs(width: int, height: int, blk_dim: int, extra_span: int) -> List[int]:
    """Pixel-compares of each block row: the separable count
    (sum_x w*ncx) * (h*ncy) with clamped candidate ranges."""
    def axis(n):
        out = []
        for p in range(0, n, blk_dim):
            e = min(blk_dim, n - p)
            lo = max(0, p - extra_span)
            hi = min(n - 1, p + e - 1 + extra_span)
            out.append(e * (hi - e + 1 - lo + 1))
        return out
    sx = sum(axis(width))
    return [sx * c for c in axis(height)]


def all_band_rows(nby: int, world: int, weights: Optional[Sequence[float]] = None) -> List[Tuple[int, int]]:
    """The bands of every rank (each rank can compute all of them: no collective needed)."""
    return [band_rows(nby, world, r, weights) for r in range(world)]


_SPANS_CACHE = {}


def balanced_spans(est, world: int, balance: bool = True) -> List[Tuple[int, int]]:
    """Bands of all ranks for an estimator's geometry (cached: this sits on the launch path)."""
    key = (est.width, est.height, est.blk_dim, est.extra_span, world, balance)
    if key not in _SPANS_CACHE:
        weights = row_costs(est.width, est.height, est.blk_dim, est.extra_span) if balance else None
        _SPANS_CACHE[key] = all_band_rows(est.blocks_y, world, weights)
    return _SPANS_CACHE[key]


def gather_bands(local: "torch.Tensor", rows: Tuple[int, int], nby: int, nbx: int, group=None,
                 spans: Optional[List[Tuple[int, int]]] = None):
    """All-gather the per-band slices of a field with ONE collective.

    `local` is this rank's (..., (rows[1]-rows[0]) * nbx) tensor (any leading dims, any dtype,
    CUDA for NCCL, CPU for gloo).  Returns the full (..., nby * nbx) field on every rank.
    Bands may have different heights, so slices are padded to the tallest band for the
    collective and trimmed afterwards.  `spans` = the bands of all ranks when the caller already
    knows them (they follow from `band_rows`); otherwise they are exchanged first.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if spans is None:
        mine = torch.tensor([rows[0], rows[1]], dtype=torch.int64, device=local.device)
        all_rows = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(all_rows, mine, group=group)
        spans = [(int(t[0]), int(t[1])) for t in all_rows]
    lead = tuple(local.shape[:-1])
    tallest = max(e - b for b, e in spans)
    padded = torch.zeros(lead + (tallest * nbx,), dtype=local.dtype, device=local.device)
    padded[..., : local.shape[-1]] = local
    gathered = torch.empty((world,) + tuple(padded.shape), dtype=local.dtype, device=local.device)
    # equal-sized views of one contiguous buffer: a single all_gather on NCCL and on gloo
    dist.all_gather(list(gathered.unbind(0)), padded.contiguous(), group=group)
    full = torch.zeros(lead + (nby * nbx,), dtype=local.dtype, device=local.device)
    covered = 0
    for r, (b, e) in enumerate(spans):
        full[..., b * nbx: e * nbx] = gathered[r][..., : (e - b) * nbx]
        covered += e - b
    if covered != nby:
        raise RuntimeError("bands do not tile the frame: %r" % (spans,))
    return full


def search_banded(est, d_cur, d_ref, pitch: int, pair_stride: int, npairs: int, group=None, stream: int = 0,
                  balance: bool = True):
    """One frame (or batch) split by block-row bands over the ranks of `group`.
    Every rank holds the full frames on its GPU, searches its band (the kernel writes the four
    output arrays into one packed int32 tensor) and ONE all_gather completes the field.
    Returns dict(mvx, mvy, ssd, score) of full (npairs, num_blocks) CUDA tensors."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    spans = balanced_spans(est, world, balance)
    b0, b1 = spans[rank]
    nb, nbx = est.num_blocks, est.blocks_x
    packed = torch.zeros((4, npairs, nb), dtype=torch.int32, device=d_cur.device)  # mvx, mvy, ssd, score bits
    est.search_device(d_cur, d_ref, pitch, pair_stride, npairs, packed[0], packed[1], packed[2], packed[3],
                      stream, b0, b1)
    full = gather_bands(packed[:, :, b0 * nbx: b1 * nbx], (b0, b1), est.blocks_y, nbx, group, spans)
    return {"mvx": full[0], "mvy": full[1], "ssd": full[2], "score": full[3].view(torch.float32)}


class _RawCuda:
    """Minimal ``__cuda_array_interface__`` view of library-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2}


class PeerField:
    """The motion field of a batch, one copy per rank, every copy mapped into every rank.

    Layout of a copy: int32 mvx | int32 mvy | uint32 ssd | float score, each ``npairs * num_blocks``
    entries, then ``ME_B200_MAX_PEERS`` uint32 barrier flags.  The memory comes from
    ``me_b200_device_alloc`` (exportable), the handles travel through ``all_gather_object``.

    The field is DOUBLE-BUFFERED (``copies`` = 2 by default): search k stores into buffer
    ``k % copies`` of every rank.  One barrier per search orders "all bands have landed" but not
    "everybody has finished reading": a rank that has passed barrier k may start search k+1 at
    once and its stores go straight into its peers' memory.  With two buffers those stores hit the
    buffer the peers read two searches ago; a peer can only be that far behind if its own search
    k (which it enqueues after its reads of search k-1, in stream order) has not run, and then the
    fast rank is still waiting in barrier k.  So the views returned for search k stay valid until
    this rank enqueues search k+2, provided its reads are ordered before its next search on the
    same stream (or the host synchronised in between)."""

    def __init__(self, est, npairs: int, group=None, copies: int = 2):
        import torch.distributed as dist
        from .lib import Field, ME_B200_MAX_PEERS
        self.est, self.npairs, self.group = est, npairs, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > ME_B200_MAX_PEERS:
            raise ValueError("at most %d ranks" % ME_B200_MAX_PEERS)
        if copies < 1:
            raise ValueError("copies >= 1")
        self.copies = copies
        self.entries = npairs * est.num_blocks
        self.copy_bytes = (4 * 4 * self.entries + 255) & ~255
        self.nbytes = self.copy_bytes * copies + 256        # buffers, then the flags
        self.base = est.device_alloc(self.nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, est.ipc_export(self.base), group=group)
        self.bases, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.bases.append(self.base)
            else:
                p = est.ipc_open(h)
                self._opened.append(p)
                self.bases.append(p)
        a = 4 * self.entries
        # fields[c][r] = buffer c of rank r
        self.fields = [[Field(b + c * self.copy_bytes, b + c * self.copy_bytes + a, b + c * self.copy_bytes + 2 * a,
                              b + c * self.copy_bytes + 3 * a) for b in self.bases] for c in range(copies)]
        self.flag_ptrs = [b + self.copy_bytes * copies for b in self.bases]
        self._tensors = [None] * copies
        self.epoch = 0
        dist.barrier(group=group)    # every rank has opened every copy before anyone writes

    def local_field(self, c: int):
        return self.fields[c][self.rank]

    def peer_fields(self, c: int):
        return [self.fields[c][r] for r in range(self.world) if r != self.rank]

    def tensors(self, c: Optional[int] = None):
        """Buffer `c` (default: the one the last search filled) of this rank as torch tensors (views)."""
        import torch
        if c is None:
            c = (self.epoch - 1) % self.copies if self.epoch else 0
        if self._tensors[c] is None:
            shape = (self.npairs, self.est.num_blocks)
            a = 4 * self.entries
            b = self.base + c * self.copy_bytes
            t = [torch.as_tensor(_RawCuda(b + i * a, shape, "<i4"), device="cuda") for i in range(3)]
            sc = torch.as_tensor(_RawCuda(b + 3 * a, shape, "<f4"), device="cuda")
            self._tensors[c] = {"mvx": t[0], "mvy": t[1], "ssd": t[2], "score": sc}
        return self._tensors[c]

    def check(self):
        """Synchronises the device and raises if a barrier of this context gave up waiting for a peer
        (the field of that search is incomplete)."""
        if self.est.peer_barrier_timed_out():
            raise RuntimeError("peer barrier timed out: a rank did not deliver its band; the field is incomplete")

    def close(self):
        import torch.distributed as dist
        if self.base:
            self._tensors = [None] * self.copies
            dist.barrier(group=self.group)   # nobody closes while a peer may still write
            for p in self._opened:
                self.est.ipc_close(p)
            self._opened = []
            self.est.device_free(self.base)
            self.base = 0


def search_banded_peer(est, field: PeerField, d_cur, d_ref, pitch: int, pair_stride: int, npairs: int,
                       stream: int = 0, balance: bool = True, timeout_ms: int = 2000, check: bool = True):
    """As :func:`search_banded`, but with no collective: the rank's band is stored into every
    rank's copy of `field` by the search itself, then one device-side barrier (flags in the
    peer-mapped memory) tells each rank that all bands have landed.  Returns views of the local
    buffer this search filled (see :class:`PeerField` for how long they stay valid).

    `check` (default) synchronises and raises when the barrier timed out, i.e. when a peer was too
    slow or failed and the field is incomplete.  A latency-critical caller passes ``check=False``
    and calls ``field.check()`` before it trusts the result."""
    spans = balanced_spans(est, field.world, balance)
    b0, b1 = spans[field.rank]
    c = field.epoch % field.copies
    est.search_device_band_peers(d_cur, d_ref, pitch, pair_stride, npairs, b0, b1, field.local_field(c),
                                 field.peer_fields(c), stream)
    field.epoch += 1
    est.peer_barrier(field.flag_ptrs, field.rank, field.epoch, timeout_ms, stream)
    if check:
        field.check()
    return field.tensors(c)
