// me_peer.cu -- peer-memory plumbing of the band sharding.
//
// One very large frame is split by block-row bands over the GPUs of one box, one process per
// GPU.  Instead of gathering the per-band slices of the motion field with a collective after the
// search, every rank maps the field buffers of all peers (CUDA IPC over NVLink/NVSwitch) and the
// search kernel itself stores each finished block into every peer's copy (me_tiled.cu, publish
// step).  What is left here:
//   peer_scatter_kernel  the same stores for block rows that a kernel without peer outputs
//                        produced (generic / small-span / SSIM / fast-search rows);
//   peer_barrier_kernel  a device-side barrier between the GPUs: after this rank's stores are
//                        fenced system-wide, lane q raises flag[my_rank] in peer q's flag array and
//                        then waits until peer q has raised its flag here.  Each GPU runs its own
//                        copy of the kernel (never two on one GPU), with a time-out.
// The reference has no multi-device code at all; this is the B200 way of
// completing a field that was computed in bands.
#include <cuda_runtime.h>
#include <stdint.h>

#include "me_device.cuh"

namespace me {

namespace {

constexpr int kMaxPeers = 8;

struct PeerOuts {
  Out o[kMaxPeers];
  int n;
};

__global__ void __launch_bounds__(256)
peer_scatter_kernel(Out local, PeerOuts peers, int nb, int first, int count) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  const size_t oi = (size_t)blockIdx.y * nb + first + i;
  int mvx = 0, mvy = 0;
  uint32_t ssd = 0;
  float score = 0.0f;
  if (local.mvx) mvx = local.mvx[oi];
  if (local.mvy) mvy = local.mvy[oi];
  if (local.ssd) ssd = local.ssd[oi];
  if (local.score) score = local.score[oi];
  for (int q = 0; q < peers.n; q++) {
    const Out &po = peers.o[q];
    if (po.mvx && local.mvx) po.mvx[oi] = mvx;
    if (po.mvy && local.mvy) po.mvy[oi] = mvy;
    if (po.ssd && local.ssd) po.ssd[oi] = ssd;
    if (po.score && local.score) po.score[oi] = score;
  }
}

struct PeerFlags {
  uint32_t *f[kMaxPeers];
};

__global__ void peer_barrier_kernel(PeerFlags flags, int npeers, int my_rank, uint32_t epoch,
                                    unsigned long long timeout_ns, int *status) {
  const int q = threadIdx.x;
  if (q >= npeers) return;
  // everything this GPU stored before this kernel (earlier kernels of the stream) is ordered
  // before the flag for every observer in the system
  __threadfence_system();
  volatile uint32_t *theirs = flags.f[q] + my_rank;
  *theirs = epoch;
  __threadfence_system();
  volatile uint32_t *mine = flags.f[my_rank] + q;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(*mine - epoch) < 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {
      if (status) atomicExch(status, 1);
      break;
    }
  }
  __threadfence_system();
}

}  // namespace

cudaError_t launch_peer_scatter(const Geom &g, int npairs, int by_begin, int by_end, const Out &local,
                                const Out *peers, int npeers, cudaStream_t s) {
  if (by_end <= by_begin || npairs <= 0 || npeers <= 0) return cudaSuccess;
  if (npeers > kMaxPeers) return cudaErrorInvalidValue;
  PeerOuts po;
  po.n = npeers;
  for (int i = 0; i < npeers; i++) po.o[i] = peers[i];
  const int first = by_begin * g.nbx, count = (by_end - by_begin) * g.nbx;
  for (int done = 0; done < npairs; done += 65535) {
    const int np = npairs - done > 65535 ? 65535 : npairs - done;
    Out lo = local;
    PeerOuts pq = po;
    const size_t off = (size_t)done * g.nbx * g.nby;
    auto shift = [&](Out &o) {
      if (o.mvx) o.mvx += off;
      if (o.mvy) o.mvy += off;
      if (o.ssd) o.ssd += off;
      if (o.score) o.score += off;
    };
    shift(lo);
    for (int i = 0; i < npeers; i++) shift(pq.o[i]);
    dim3 grid((unsigned)((count + 255) / 256), (unsigned)np);
    peer_scatter_kernel<<<grid, 256, 0, s>>>(lo, pq, g.nbx * g.nby, first, count);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_peer_barrier(uint32_t *const *flags, int npeers, int my_rank, uint32_t epoch,
                                unsigned long long timeout_ns, int *d_status, cudaStream_t s) {
  if (npeers < 1 || npeers > kMaxPeers || my_rank < 0 || my_rank >= npeers) return cudaErrorInvalidValue;
  PeerFlags pf;
  for (int i = 0; i < kMaxPeers; i++) pf.f[i] = i < npeers ? flags[i] : nullptr;
  peer_barrier_kernel<<<1, 32, 0, s>>>(pf, npeers, my_rank, epoch, timeout_ns, d_status);
  return cudaGetLastError();
}

}  // namespace me
