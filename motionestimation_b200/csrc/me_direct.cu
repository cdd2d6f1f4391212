// me_direct.cu -- small-span full search (extra span R <= 4, i.e. at most 81 candidates per
// block), the memory-bound end of the path (only +-0..+-2 ranges are
// bandwidth-bound with u8 frames).
//
// Reference being replaced: the same main.c:18-82 scan as the tuned kernel; with so few
// candidates the rotating-accumulator streaming of me_tiled.cu cannot amortise its set-up, so
// this kernel maps B lanes to one (block, candidate) pair instead:
//   * a CTA owns a 128 x 32 pixel tile of blocks (8 x 2 blocks of 16x16 or 16 x 4 blocks of 8x8);
//     the current tile and the reference tile + R halo are staged in shared memory once with
//     coalesced 16-byte loads (each frame byte is read from HBM ~1.1 times);
//   * a group of B lanes scores one (block, candidate) unit, one block row per lane: the aligned
//     reference words around the candidate column are funnel-shifted into place and compared with
//     VABSDIFF4.U8 + IDP.4A.U8.U8 (exact integer SSD), the rows are added with shuffles;
//   * key = ssd << 8 | raster index of the candidate in the (2R+1)^2 grid, one 32-bit shared
//     atomicMin per candidate; the unsigned minimum is the reference's first strict minimum in
//     y-major/x-minor order (main.c:53-62) because clamped-away candidates are simply skipped.
#include <stdlib.h>

#include "me_device.cuh"

namespace me {

namespace {

constexpr int kTX = 128, kTY = 32;  // tile of pixels per CTA
constexpr int kMaxR = 4;
constexpr int kRefStage = kTX + 32;                 // staged bytes per reference row: >= 15 + kTX + 2R + 3, multiple of 16
// Row pitches in 32-bit words are ODD (41, 33): the B lanes of a unit read B different rows at the
// same column, and an odd word stride spreads them over B different banks.
constexpr int kRefPitchW = kRefStage / 4 + 1;
constexpr int kCurPitchW = kTX / 4 + 1;
constexpr int kRefRows = kTY + 2 * kMaxR;
constexpr int kThreads = 256;

template <int B, bool ROWSPLIT>
__global__ void __launch_bounds__(kThreads)
direct_search_kernel(Geom g, Frames f, Out o) {
  constexpr int NBX = kTX / B, NBY = kTY / B, NBLK = NBX * NBY, WPR = B / 4;
  __shared__ uint32_t s_cur[kTY * kCurPitchW];
  __shared__ uint32_t s_ref[kRefRows * kRefPitchW];
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
  // reference columns start at the 16-aligned column left of tx0 - R
  const int rx0 = (tx0 - R) & ~15;  // may be negative
  const int ex = tx0 - R - rx0;     // 0..15 bytes between the aligned origin and tx0 - R
  // all global loads of a thread are issued before the first shared store, so their latencies
  // overlap: one 16-byte piece of the current tile, up to two of the reference tile
  static_assert(kTY * (kTX / 16) == kThreads, "one current-tile piece per thread");
  constexpr int kRefPieces = (kRefRows * (kRefStage / 16) + kThreads - 1) / kThreads;
  const int nref = (kTY + 2 * R) * (kRefStage / 16);
  uint4 vc, vr[kRefPieces];
  {
    const int r = threadIdx.x / (kTX / 16), k = threadIdx.x - r * (kTX / 16);
    vc = load16(cur, tx0 + 16 * k, ty0 + r);
  }
#pragma unroll
  for (int j = 0; j < kRefPieces; j++) {
    const int i = threadIdx.x + j * kThreads;
    const int r = i / (kRefStage / 16), k = i - r * (kRefStage / 16);
    vr[j] = i < nref ? load16(ref, rx0 + 16 * k, ty0 - R + r) : make_uint4(0u, 0u, 0u, 0u);
  }
  {
    const int r = threadIdx.x / (kTX / 16), k = threadIdx.x - r * (kTX / 16);
    uint32_t *d = s_cur + r * kCurPitchW + 4 * k;
    d[0] = vc.x; d[1] = vc.y; d[2] = vc.z; d[3] = vc.w;
  }
#pragma unroll
  for (int j = 0; j < kRefPieces; j++) {
    const int i = threadIdx.x + j * kThreads;
    if (i < nref) {
      const int r = i / (kRefStage / 16), k = i - r * (kRefStage / 16);
      uint32_t *d = s_ref + r * kRefPitchW + 4 * k;
      d[0] = vr[j].x; d[1] = vr[j].y; d[2] = vr[j].z; d[3] = vr[j].w;
    }
  }
  if (threadIdx.x < NBLK) s_best[threadIdx.x] = 0xffffffffu;
  __syncthreads();

  if constexpr (ROWSPLIT) {
    // One (block, candidate) UNIT per group of B lanes, one block row per lane: with only 1..81
    // candidates per block a whole-candidate-per-thread mapping leaves most of the CTA idle behind a
    // few long dependent chains (at +-0 just 16 of 256 threads had work); split by rows, every lane
    // does 5 LDS + 4 x (SHF, VABSDIFF4, IDP.4A), the group adds up with log2(B) shuffles and its
    // first lane does the shared atomicMin.  Passes are uniform over the CTA, so the shuffles always
    // run with the full mask.
    constexpr int UPP = kThreads / B;                        // units per pass
    const int lane_row = threadIdx.x % B, unit_in_pass = threadIdx.x / B;
    const int total_units = NBLK * ncand;
    for (int base = 0; base < total_units; base += UPP) {
      const int unit = base + unit_in_pass;
      uint32_t ssd = 0;
      bool valid = false;
      int blk = 0, c = 0;
      if (unit < total_units) {
        blk = unit / ncand;
        c = unit - blk * ncand;
        const int by_ = blk / NBX, bx_ = blk - by_ * NBX;
        const int dyi = c / nd, dxi = c - dyi * nd;          // window-relative offsets, mv = d - R
        const int x0 = tx0 + bx_ * B, y0 = ty0 + by_ * B;    // block origin in the frame
        if (x0 < g.W && y0 < y_end) {
          const int w = min(B, g.W - x0), h = min(B, g.H - y0);
          // clamped window (main.c:73-76): the candidate must lie inside the frame
          const int cx = x0 + dxi - R, cy = y0 + dyi - R;
          valid = !(cx < 0 || cy < 0 || cx + w > g.W || cy + h > g.H);
          if (valid && lane_row < h) {
            const int u = ex + bx_ * B + dxi;                // byte column in s_ref rows
            const uint32_t shift = 8u * (uint32_t)(u & 3);
            const uint32_t *rp = s_ref + (by_ * B + dyi + lane_row) * kRefPitchW + (u >> 2);
            const uint32_t *cp = s_cur + (by_ * B + lane_row) * kCurPitchW + bx_ * WPR;
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
          }
        }
      }
  #pragma unroll
      for (int off = B / 2; off; off >>= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, off);
      if (valid && lane_row == 0) atomicMin(&s_best[blk], (ssd << 8) | (uint32_t)c);
    }
  } else {
  // one (block, candidate) pair per thread-iteration; candidates vary fastest (consecutive lanes
    // read the same rows: broadcasts, no bank conflicts)
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
      const uint32_t *rp = s_ref + (by_ * B + dyi) * kRefPitchW + (u >> 2);
      const uint32_t *cp = s_cur + (by_ * B) * kCurPitchW + bx_ * WPR;
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
        rp += kRefPitchW;
        cp += kCurPitchW;
      }
      atomicMin(&s_best[blk], (ssd << 8) | (uint32_t)c);
    }
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

// ---- R = 0: the purely memory-bound end.  The only candidate of a block is the co-located one
// (main.c:73-76 clamp the window to the block itself), so the search is one streaming pass over
// both frames straight from global memory (aligned: dx = 0), VABSDIFF4 + IDP.4A.  A warp owns the
// 128 / B blocks that span 128 contiguous bytes of one block row: every load instruction of the
// warp fetches whole 128-byte lines (LPR lanes per row, 32 / LPR rows per instruction), all
// B / (32 / LPR) = 4 loads per frame of a lane are issued before the first is consumed, and the
// lanes that hold the rows of one block are added with two or one shuffles.  ~60 instructions per
// 128 input bytes of a lane: the kernel sits on the HBM roof, not on the issue rate like the
// staged kernel above.
template <int B, int LPR>
__global__ void __launch_bounds__(256)
zero_span_kernel(Geom g, Frames f, Out o, int nbx_groups) {
  // LPR = lanes (= blocks) per row piece of a warp: LPR * B contiguous bytes per row
  constexpr int RPI = 32 / LPR;                     // rows per load instruction: 4 or 2
  constexpr int NIT = B / RPI;                      // load instructions per frame: 4
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int bc = lane % LPR, rq = lane / LPR;
  const int wx = warp % nbx_groups, wy = warp / nbx_groups;     // warp -> (group of LPR blocks, block row)
  if (wy >= g.by_count) return;
  const int by = g.by_begin + wy;
  const int bx = wx * LPR + bc;
  const int x0 = bx * B, y0 = by * B;
  const bool blk_ok = bx < g.nbx;
  const int w = blk_ok ? min(B, g.W - x0) : 0, h = min(B, g.H - y0);
  const size_t base = (size_t)blockIdx.y * f.pair_stride + (size_t)y0 * f.pitch + x0;
  uint32_t c[NIT][B / 4], r[NIT][B / 4];
#pragma unroll
  for (int i = 0; i < NIT; i++) {
    const int row = i * RPI + rq;
    const size_t off = base + (size_t)row * f.pitch;
#pragma unroll
    for (int k = 0; k < B / 4; k++) c[i][k] = r[i][k] = 0u;
    if (w == B && row < h) {
      if constexpr (B == 16) {
        const uint4 cv = *reinterpret_cast<const uint4 *>(f.cur + off), rv = *reinterpret_cast<const uint4 *>(f.ref + off);
        c[i][0] = cv.x; c[i][1] = cv.y; c[i][2] = cv.z; c[i][3] = cv.w;
        r[i][0] = rv.x; r[i][1] = rv.y; r[i][2] = rv.z; r[i][3] = rv.w;
      } else {
        const uint2 cv = *reinterpret_cast<const uint2 *>(f.cur + off), rv = *reinterpret_cast<const uint2 *>(f.ref + off);
        c[i][0] = cv.x; c[i][1] = cv.y;
        r[i][0] = rv.x; r[i][1] = rv.y;
      }
    } else if (w > 0 && row < h) {  // partial-width block at the right frame edge: columns >= w do not exist
      for (int b = 0; b < w; b++) {
        c[i][b >> 2] |= (uint32_t)f.cur[off + b] << (8 * (b & 3));
        r[i][b >> 2] |= (uint32_t)f.ref[off + b] << (8 * (b & 3));
      }
    }
  }
  uint32_t ssd = 0;
#pragma unroll
  for (int i = 0; i < NIT; i++)
#pragma unroll
    for (int k = 0; k < B / 4; k++) {
      const uint32_t d = __vabsdiffu4(c[i][k], r[i][k]);   // rows that do not exist are 0 vs 0
      ssd = __dp4a(d, d, ssd);
    }
#pragma unroll
  for (int s_ = LPR; s_ < 32; s_ <<= 1) ssd += __shfl_xor_sync(0xffffffffu, ssd, s_);
  if (blk_ok && rq == 0) {
    const size_t oi = (size_t)blockIdx.y * g.nbx * g.nby + (size_t)by * g.nbx + bx;
    if (o.mvx) o.mvx[oi] = 0;      // main.c:58-59: the only candidate is the block's own position
    if (o.mvy) o.mvy[oi] = 0;
    if (o.ssd) o.ssd[oi] = ssd;
    if (o.score) o.score[oi] = __fdiv_rn((float)ssd, (float)(w * h));  // main.c:27
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
    // R = 0 on a 16-byte aligned layout: the streaming kernel
    const bool stream_ok = g.R == 0 && (f.pitch & 15) == 0 && (f.pair_stride & 15) == 0 &&
                           ((((uintptr_t)f.cur) | ((uintptr_t)f.ref)) & 15) == 0;
    if (stream_ok) {
      // blocks of one warp (LPR), measured on 4K / 1080p batches: 16x16: 2 (32-byte row pieces, all 16
      // rows in one load instruction: 4.9 TB/s; 4 or 8 blocks: 2.8-3.0 TB/s); 8x8: 8 (64-byte pieces,
      // 4 rows per instruction: 3.6 TB/s; 4 or 16 blocks: 2.7 TB/s)
      int bpw = g.B == 16 ? 2 : 8;
      if (const char *e = getenv("ME_B200_R0_LPR")) bpw = atoi(e);
      const int nbx_groups = (g.nbx + bpw - 1) / bpw;
      const long long warps = (long long)nbx_groups * g.by_count;
      dim3 zg((unsigned)((warps + 7) / 8), (unsigned)n);
      if (g.B == 16) {
        if (bpw == 8) zero_span_kernel<16, 8><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
        else if (bpw == 4) zero_span_kernel<16, 4><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
        else zero_span_kernel<16, 2><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
      } else {
        if (bpw == 4) zero_span_kernel<8, 4><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
        else if (bpw == 8) zero_span_kernel<8, 8><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
        else zero_span_kernel<8, 16><<<zg, 256, 0, s>>>(g, ff, oo, nbx_groups);
      }
    } else if (g.R == 0) {
      if (g.B == 16) direct_search_kernel<16, true><<<grid, kThreads, 0, s>>>(g, ff, oo);
      else direct_search_kernel<8, true><<<grid, kThreads, 0, s>>>(g, ff, oo);
    } else {
      if (g.B == 16) direct_search_kernel<16, false><<<grid, kThreads, 0, s>>>(g, ff, oo);
      else direct_search_kernel<8, false><<<grid, kThreads, 0, s>>>(g, ff, oo);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    done += n;
  }
  return cudaSuccess;
}

}  // namespace me
