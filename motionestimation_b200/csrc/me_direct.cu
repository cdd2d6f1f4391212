// me_direct.cu -- small-span full search (extra span R <= 4, i.e. at most 81 candidates per
// block), the memory-bound end of the path (only +-0..+-2 ranges are
// bandwidth-bound with u8 frames).
//
// Reference being replaced: the same main.c:18-82 scan as the tuned kernel; with so few
// candidates the rotating-accumulator streaming of me_tiled.cu cannot amortise its set-up, so
// this kernel maps one thread to one (block, candidate) pair instead:
//   * a CTA owns a 128 x 32 pixel tile of blocks (8 x 2 blocks of 16x16 or 16 x 4 blocks of 8x8);
//     the current tile and the reference tile + R halo are staged in shared memory once with
//     coalesced 16-byte loads (each frame byte is read from HBM ~1.1 times);
//   * a thread scores candidates idx = tid, tid + 256, ... of the tile: per row, the aligned
//     reference words around the candidate column are funnel-shifted into place and compared with
//     VABSDIFF4.U8 + IDP.4A.U8.U8 (exact integer SSD);
//   * key = ssd << 8 | raster index of the candidate in the (2R+1)^2 grid, one 32-bit shared
//     atomicMin per candidate; the unsigned minimum is the reference's first strict minimum in
//     y-major/x-minor order (main.c:53-62) because clamped-away candidates are simply skipped.
#include "me_device.cuh"

namespace me {

namespace {

constexpr int kTX = 128, kTY = 32;  // tile of pixels per CTA
constexpr int kMaxR = 4;
constexpr int kRefPitch = kTX + 32;                 // >= 15 + kTX + 2R + 3, multiple of 16
constexpr int kRefRows = kTY + 2 * kMaxR;
constexpr int kThreads = 256;

template <int B>
__global__ void __launch_bounds__(kThreads)
direct_search_kernel(Geom g, Frames f, Out o) {
  constexpr int NBX = kTX / B, NBY = kTY / B, NBLK = NBX * NBY, WPR = B / 4;
  __shared__ __align__(16) uint8_t s_cur[kTY * kTX];
  __shared__ __align__(16) uint8_t s_ref[kRefRows * kRefPitch];
  __shared__ uint32_t s_best[NBLK];

  const int R = g.R, nd = 2 * R + 1, ncand = nd * nd;
  const int tx0 = blockIdx.x * kTX;                       // tile origin in pixels
  const int ty0 = (g.by_begin * g.B) + blockIdx.y * kTY;
  const int y_end = min(g.H, (g.by_begin + g.by_count) * g.B);  // rows of this launch's band
  const uint8_t *cur = f.cur + (size_t)blockIdx.z * f.pair_stride;
  const uint8_t *ref = f.ref + (size_t)blockIdx.z * f.pair_stride;

  // stage the current tile and the reference tile + halo (zeros outside the frame); 16-byte loads
  // wherever the 16 bytes lie inside the frame and the layout is 16-byte aligned
  const bool vec = ((f.pitch & 15) == 0) && ((((uintptr_t)cur | (uintptr_t)ref) & 15) == 0);
  auto load16 = [&](const uint8_t *base, int x, int y) -> uint4 {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y < 0 || y >= g.H || x + 15 < 0 || x >= g.W) return v;
    const uint8_t *q = base + (size_t)y * f.pitch;
    if (vec && x >= 0 && x + 16 <= g.W) return *reinterpret_cast<const uint4 *>(q + x);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int b = 0; b < 16; b++)
      if (x + b >= 0 && x + b < g.W) w[b >> 2] |= (uint32_t)q[x + b] << (8 * (b & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
  };
  for (int i = threadIdx.x; i < kTY * (kTX / 16); i += kThreads) {
    const int r = i / (kTX / 16), k = i - r * (kTX / 16);
    reinterpret_cast<uint4 *>(s_cur)[i] = load16(cur, tx0 + 16 * k, ty0 + r);
  }
  // reference columns start at the 16-aligned column left of tx0 - R
  const int rx0 = (tx0 - R) & ~15;  // may be negative
  const int ex = tx0 - R - rx0;     // 0..15 bytes between the aligned origin and tx0 - R
  for (int i = threadIdx.x; i < (kTY + 2 * R) * (kRefPitch / 16); i += kThreads) {
    const int r = i / (kRefPitch / 16), k = i - r * (kRefPitch / 16);
    reinterpret_cast<uint4 *>(s_ref)[r * (kRefPitch / 16) + k] = load16(ref, rx0 + 16 * k, ty0 - R + r);
  }
  if (threadIdx.x < NBLK) s_best[threadIdx.x] = 0xffffffffu;
  __syncthreads();

  // one (block, candidate) pair per thread-iteration; candidates vary fastest
  for (int idx = threadIdx.x; idx < NBLK * ncand; idx += kThreads) {
    const int blk = idx / ncand, c = idx - blk * ncand;
    const int by_ = blk / NBX, bx_ = blk - by_ * NBX;
    const int dyi = c / nd, dxi = c - dyi * nd;          // window-relative offsets, mv = d - R
    const int x0 = tx0 + bx_ * B, y0 = ty0 + by_ * B;    // block origin in the frame
    if (x0 >= g.W || y0 >= y_end) continue;
    const int w = min(B, g.W - x0), h = min(B, g.H - y0);
    // clamped window (main.c:73-76): the candidate must lie inside the frame
    const int cx = x0 + dxi - R, cy = y0 + dyi - R;
    if (cx < 0 || cy < 0 || cx + w > g.W || cy + h > g.H) continue;
    const int u = ex + bx_ * B + dxi;                    // byte column in s_ref rows
    const uint32_t shift = 8u * (uint32_t)(u & 3);
    const uint32_t *rp = reinterpret_cast<const uint32_t *>(s_ref) + (by_ * B + dyi) * (kRefPitch / 4) + (u >> 2);
    const uint32_t *cp = reinterpret_cast<const uint32_t *>(s_cur) + (by_ * B) * (kTX / 4) + bx_ * WPR;
    uint32_t ssd = 0;
    for (int r = 0; r < h; r++) {
      uint32_t raw[WPR + 1];
#pragma unroll
      for (int k = 0; k <= WPR; k++) raw[k] = rp[k];
#pragma unroll
      for (int k = 0; k < WPR; k++) {
        uint32_t rv = __funnelshift_r(raw[k], raw[k + 1], shift);
        uint32_t cv = cp[k];
        // partial-width blocks: compare only the w valid columns (both zero-padded otherwise,
        // but the reference side holds real pixels there)
        if (w < B) {
          const int left = w - 4 * k;
          const uint32_t m = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : (0xffffffffu >> (8 * (4 - left))));
          rv &= m;
          cv &= m;
        }
        const uint32_t d = __vabsdiffu4(cv, rv);
        ssd = __dp4a(d, d, ssd);
      }
      rp += kRefPitch / 4;
      cp += kTX / 4;
    }
    atomicMin(&s_best[blk], (ssd << 8) | (uint32_t)c);
  }
  __syncthreads();

  if (threadIdx.x < NBLK) {
    const int by_ = threadIdx.x / NBX, bx_ = threadIdx.x - by_ * NBX;
    const int x0 = tx0 + bx_ * B, y0 = ty0 + by_ * B;
    if (x0 < g.W && y0 < y_end) {
      const uint32_t key = s_best[threadIdx.x];
      const int c = (int)(key & 0xffu), dyi = c / nd, dxi = c - dyi * nd;
      const uint32_t ssd = key >> 8;
      const int w = min(B, g.W - x0), h = min(B, g.H - y0);
      const size_t oi = (size_t)blockIdx.z * g.nbx * g.nby + (size_t)(y0 / B) * g.nbx + x0 / B;
      if (o.mvx) o.mvx[oi] = dxi - R;   // main.c:58
      if (o.mvy) o.mvy[oi] = dyi - R;   // main.c:59
      if (o.ssd) o.ssd[oi] = ssd;
      if (o.score) o.score[oi] = __fdiv_rn((float)ssd, (float)(w * h));  // main.c:27
    }
  }
}

}  // namespace

bool direct_supported(const Geom &g, size_t pitch, size_t pair_stride, const void *cur, const void *ref) {
  if (g.B != 8 && g.B != 16) return false;
  if (g.R < 0 || g.R > kMaxR) return false;
  if ((pitch & 3) || (pair_stride & 3)) return false;  // 32-bit row loads
  if (((uintptr_t)cur & 3) || ((uintptr_t)ref & 3)) return false;
  return true;
}

cudaError_t launch_direct(const Geom &g, const Frames &f, int npairs, const Out &o, cudaStream_t s) {
  const int rows_px = g.by_count * g.B;
  int done = 0;
  while (done < npairs) {
    const int n = npairs - done > 65535 ? 65535 : npairs - done;
    Frames ff = f;
    ff.cur += (size_t)done * f.pair_stride;
    ff.ref += (size_t)done * f.pair_stride;
    Out oo = o;
    const size_t off = (size_t)done * g.nbx * g.nby;
    if (oo.mvx) oo.mvx += off;
    if (oo.mvy) oo.mvy += off;
    if (oo.ssd) oo.ssd += off;
    if (oo.score) oo.score += off;
    dim3 grid((g.W + kTX - 1) / kTX, (rows_px + kTY - 1) / kTY, n);
    if (g.B == 16) direct_search_kernel<16><<<grid, kThreads, 0, s>>>(g, ff, oo);
    else direct_search_kernel<8><<<grid, kThreads, 0, s>>>(g, ff, oo);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    done += n;
  }
  return cudaSuccess;
}

}  // namespace me
