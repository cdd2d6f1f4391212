// me_direct.cu -- small-span full search (extra span R <= 4, i.e. at most 81 candidates per
// block), the memory-bound end of the path (only +-0..+-2 ranges are
// bandwidth-bound with u8 frames).
//
// Reference being replaced: the same main.c:18-82 scan as the tuned kernel.  With so few candidates
// the rotating-accumulator streaming of me_tiled.cu cannot amortise its per-task set-up, so three
// dedicated kernels run here (launch_direct picks):
//   * zero_span_kernel          R = 0: one streaming pass over both frames from global memory;
//   * stream_search_kernel      1 <= R <= 4, 16x16 blocks, width a multiple of 16: a warp walks down a
//                               stripe of the frame, rows arrive by TMA, every (dx, dy) of a block column is
//                               a live accumulator in registers -- no per-candidate set-up at all;
//   * direct_search_kernel      everything else (8x8 blocks, other widths): a CTA stages a 128 x 32 pixel
//                               tile + halo in shared memory, a group of B lanes (or one thread) scores one
//                               (block, candidate) with VABSDIFF4.U8 + IDP.4A.U8.U8, 32-bit shared atomicMin
//                               of ssd << 8 | raster index.
// In all of them the unsigned minimum of (cost, raster index of the candidate) is the reference's first
// strict minimum in y-major/x-minor order (main.c:53-62), and clamped-away candidates (main.c:73-76) are
// simply never folded in.
#include <stdlib.h>

#include "me_device.cuh"
#include "me_tma.cuh"

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

// ---- 1 <= R <= 4, 16x16 blocks, frame width a multiple of 16: the register-streaming kernel --------
//
// With (2R+1)^2 <= 81 candidates per block a frame pair carries only 9..81 pixel-compares per
// input byte, so the search has to run at streaming speed: +-1 is bound by HBM, +-2 sits on both
// roofs at once, +-4 on the integer pipes.  The kernel has no per-candidate set-up at all:
//   * a WARP owns a vertical stripe of the frame, 32 / G block columns wide, and walks DOWN it one
//     pixel row per step.  A lane (block column bx, dx group gi) keeps K horizontal offsets
//     dx = gi*K - R + k and all 2R+1 vertical offsets of them as live accumulators.
//   * rows arrive by TMA (cp.async.bulk.tensor, u8 tensor maps over (x, y, pair)): boxes of two
//     rows, NB boxes in flight per warp, one mbarrier per box; the reference box starts 16 bytes left
//     of the warp's span and ends 16 bytes right of it, so the +-R halo is part of the same box and
//     frame borders cost nothing (out-of-frame bytes are zero-filled, out-of-frame candidates are
//     never folded in, main.c:73-76).  The warp re-arms its own ring (lane 0, every second step): no
//     producer warp, no CTA-wide synchronisation.  Each frame byte leaves HBM once (+ 2R rows per
//     stripe).
//   * step y: the reference row y is byte-aligned to each dx by funnel shifts and multiplied
//     (IDP.4A.U8.U8) with the 2R+1 current rows y-R..y+R, read back from the ring (the first rows of
//     the current ring are mirrored behind its end, so the window is one base register + immediates)
//     -- row y of the reference is row (y - dy - y0) of the candidate dy of the block that contains
//     current row y - dy.  Exact SSD by  SSD = sum cur^2 + sum ref^2 - 2 sum cur*ref : the cross
//     term is one IDP.4A per 4 pixels; sum ref^2 of a candidate is a difference of a running prefix
//     sum of row energies (4 IDP.4A per step and dx), sum cur^2 of a block likewise (4 per step).
//     Candidates are ranked by t = sum ref^2 - 2 sum cur*ref (sum cur^2 is common to a block).
//   * a candidate finishes when its current row is the last row of its block -- at most one dy per
//     step (two next to a partial bottom block), a warp-uniform event: key = (t + 2^24) << 7 | raster
//     index of (dy, dx) folds into the lane's running minimum; the unsigned minimum is the
//     reference's first strict minimum in y-major/x-minor order (main.c:53-62).  Candidates outside
//     the clamped window and those of blocks another stripe owns are simply not folded in.  After
//     the block's last dy the G lanes of a block column combine by shuffle and one lane stores
//     MV / SSD / score.
// The loop body is one step; the IDP.4A stream on the FMA-heavy pipe is the only resource meant to
// saturate.  Per step at +-2 (K = 5): 100 cross-term + 24 energy IDP.4A.
constexpr int kStreamWarps = 1;   // one warp per CTA: everything that addresses the ring and the tensor maps is block-uniform
constexpr uint32_t kTBias = 1u << 24;
constexpr int kBoxRows = 2;

struct StreamParams {
  int nstripes;     // stripes per pair: stripe i owns block rows by_begin + [i*rows/n, (i+1)*rows/n)
  int ncg;          // column groups (= warps) per stripe
  long long total;  // warps of work in this launch
};

template <int R, int K, int G, int NB>
struct StreamLayout {
  static constexpr int ND = 2 * R + 1;
  static constexpr int CPW = 32 / G;                 // block columns per warp
  // staged reference row = TMA box width: 16 bytes left of the span (the last 4 of them: halo), own
  // bytes, 16 bytes right (the first 4: halo), padded so that a two-row box is a multiple of 128 bytes
  static constexpr int ROWB = (CPW * 16 + 32 + 63) / 64 * 64;
  static constexpr int CROWB = (CPW * 16 + 63) / 64 * 64;        // staged current row: own bytes (+ the same padding)
  static constexpr int MB = (2 * R + kBoxRows - 1) / kBoxRows;   // current boxes mirrored behind the ring's end
  static constexpr int DCB = NB + MB + 1;            // current boxes alive at once
  static constexpr int kRefBytes = NB * kBoxRows * ROWB;
  static constexpr int kCurBytes = (DCB + MB) * kBoxRows * CROWB;
  static constexpr int kWarpBytes = kRefBytes + kCurBytes;
  static_assert((kBoxRows * ROWB) % 128 == 0 && (kBoxRows * CROWB) % 128 == 0, "TMA destinations are 128-byte aligned");
};

template <int R, int K, int G, int NB, int MINB>
__global__ void __launch_bounds__(kStreamWarps * 32, MINB)
stream_search_kernel(const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_cur,
                     Geom g, Out o, StreamParams sp) {
  using L = StreamLayout<R, K, G, NB>;
  constexpr int ND = L::ND, CPW = L::CPW, ROWB = L::ROWB, CROWB = L::CROWB, DCB = L::DCB, MB = L::MB;
  static_assert(K * G >= ND && K <= 5 && R >= 1 && R <= 4 && NB >= 2, "dx groups must cover the span");
  extern __shared__ __align__(128) uint8_t stream_smem[];
  __shared__ __align__(8) uint64_t bars[kStreamWarps][NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long item = (long long)blockIdx.x * kStreamWarps + warp;
  if (item >= sp.total) return;  // the whole warp leaves together (no CTA-wide synchronisation below)
  const int per_pair = sp.nstripes * sp.ncg;
  const int pair = (int)(item / per_pair);
  const int rem = (int)(item - (long long)pair * per_pair);
  const int stripe = rem / sp.ncg, cg = rem - stripe * sp.ncg;
  const int b0 = g.by_begin + (int)((long long)g.by_count * stripe / sp.nstripes);          // block rows [b0, b1)
  const int b1 = g.by_begin + (int)((long long)g.by_count * (stripe + 1) / sp.nstripes);
  const int col_raw = lane / G, gi = lane - col_raw * G;
  const int col = col_raw < CPW ? col_raw : CPW - 1;               // (G = 3: lanes 30, 31 idle along on the last column)
  const int bx = cg * CPW + col;
  const bool active = col_raw < CPW && bx < g.nbx;
  const int x0 = bx * 16;
  const int H = g.H;
  const int y_start = b0 * 16 - R;                                 // rows above the frame: zero-filled by TMA
  const int y_end = min(H - 1, b1 * 16 - 1 + R);
  uint8_t *ring_ref = stream_smem + (size_t)warp * L::kWarpBytes;
  uint8_t *ring_cur = ring_ref + L::kRefBytes;
  uint64_t *bar = bars[warp];

  // horizontal clamp (main.c:73,75): candidate column x0 + dx must lie in [0, W - 16]
  uint32_t lvalid = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int dxr = gi * K + k - R;
    if (active && dxr <= R && x0 + dxr >= 0 && x0 + dxr + 16 <= g.W) lvalid |= 1u << k;
  }
  const uint32_t lane_idx = (uint32_t)(gi * K);  // raster index of the candidate = (dy+R) * ND + gi*K + k

  // ---- producer: box j = reference rows y_start + 2j, +1 and current rows y_start + R + 2j, +1
  const int nboxes = (y_end - y_start + kBoxRows) / kBoxRows;
  const int x_span = cg * CPW * 16;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NB; j++) mbar_init(&bar[j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t s_ref = smem_u32(ring_ref), s_cur = smem_u32(ring_cur);
  int jb_i = 0;                    // next box to load
  int jslot_i = 0;                 // its barrier
  uint32_t roff_i = 0, coff_i = 0; // its slots (byte offsets)
  auto load_box = [&]() {          // lane 0 only
    if (jb_i < nboxes) {
      const bool mirror = coff_i < (uint32_t)(MB * kBoxRows * CROWB);
      // order this warp's earlier generic-proxy reads of the slots before the async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint64_t *b = &bar[jslot_i];
      mbar_arrive_expect_tx(b, (uint32_t)(kBoxRows * ROWB + (mirror ? 2 : 1) * kBoxRows * CROWB));
      const int yy = y_start + kBoxRows * jb_i;
      // (x in 4-byte words: the maps describe the frames as u32 so that a box may be wider than 256 bytes)
      tma_load_3d_s(s_ref + roff_i, &map_ref, b, (x_span - 16) >> 2, yy, pair);
      tma_load_3d_s(s_cur + coff_i, &map_cur, b, x_span >> 2, yy + R, pair);
      if (mirror) tma_load_3d_s(s_cur + coff_i + DCB * kBoxRows * CROWB, &map_cur, b, x_span >> 2, yy + R, pair);
    }
    jb_i++;
    jslot_i = jslot_i + 1 == NB ? 0 : jslot_i + 1;
    roff_i = roff_i + kBoxRows * ROWB == (uint32_t)L::kRefBytes ? 0u : roff_i + kBoxRows * ROWB;
    coff_i = coff_i + kBoxRows * CROWB == (uint32_t)(DCB * kBoxRows * CROWB) ? 0u : coff_i + kBoxRows * CROWB;
  };
  if (lane == 0) {
#pragma unroll 1
    for (int j = 0; j < NB - 1; j++) load_box();
  }

  uint32_t acc[K][ND];    // sum cur*ref of the live candidate of every (dx, dy)
  uint32_t ss[K][ND];     // prefix sum of the reference row energies when that candidate started
  uint32_t S[K];          // running prefix sum of the reference row energies per dx
#pragma unroll
  for (int k = 0; k < K; k++) {
    S[k] = 0u;
#pragma unroll
    for (int i = 0; i < ND; i++) acc[k][i] = ss[k][i] = 0u;
  }
  uint32_t Sc = 0u, ScStart = 0u, A0 = 0u, A1 = 0u;  // sum cur^2: prefix, value at the block's start, finished blocks
  uint32_t best = 0xffffffffu;

  auto publish = [&](const int by) {
    uint32_t m = best;
    if (G == 2) m = min(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if (G == 3) {
      const int base = col_raw * 3;
      const uint32_t m0 = __shfl_sync(0xffffffffu, best, base & 31);
      const uint32_t m1 = __shfl_sync(0xffffffffu, best, (base + 1) & 31);
      const uint32_t m2 = __shfl_sync(0xffffffffu, best, (base + 2) & 31);
      m = min(m0, min(m1, m2));
    }
    if (active && gi == 0) {
      const int idx = (int)(m & 127u);
      const int dyi = idx / ND, dxi = idx - dyi * ND;
      const uint32_t a = (by & 1) ? A1 : A0;
      const uint32_t ssd = (m >> 7) - kTBias + a;       // t + sum cur^2
      const int hb = min(16, H - by * 16);
      const size_t oi = (size_t)pair * g.nbx * g.nby + (size_t)by * g.nbx + bx;
      if (o.mvx) o.mvx[oi] = dxi - R;   // main.c:58
      if (o.mvy) o.mvy[oi] = dyi - R;   // main.c:59
      if (o.ssd) o.ssd[oi] = ssd;
      if (o.score) o.score[oi] = __fdiv_rn((float)ssd, (float)(16 * hb));  // main.c:27
    }
    best = 0xffffffffu;
  };

  // the candidates dy = di - R of all dx finish at step y: fold them in, restart their accumulators;
  // returns the block row to publish (its last candidate row has just finished) or -1
  auto finish = [&](const int di, const int y) -> int {
    const int dy = di - R;
    const int c = y - dy;                       // their current row: the last row of its block
    const int by = c >> 4, y0 = by * 16;
    const int hb = min(16, H - y0);
    const bool ok = by >= b0 && by < b1 && c == y0 + hb - 1 && y0 + dy >= 0;
    const uint32_t okmask = ok ? lvalid : 0u;
#pragma unroll
    for (int d = 0; d < ND; d++) {
      if (di == d) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          const uint32_t e = S[k] - ss[k][d];                     // sum ref^2 over the candidate
          const uint32_t t = e - acc[k][d] - acc[k][d] + kTBias;  // + 2^24 keeps it positive
          const uint32_t key = (t << 7) + (uint32_t)(d * ND + k) + lane_idx;
          best = min(best, ((okmask >> k) & 1u) ? key : 0xffffffffu);
          acc[k][d] = 0u;
          ss[k][d] = S[k];
        }
      }
    }
    // vertical clamp (main.c:74,76): dy was the last candidate row of this block
    return (ok && dy == min(R, H - hb - y0)) ? by : -1;
  };

  // ---- consumer: this lane's bytes of the reference row being read / of the newest current row
  const uint8_t *rd0 = ring_ref + 16 * col, *rd = rd0;
  const uint8_t *cd0 = ring_cur + 16 * col;
  uint32_t coff = 0;     // primary offset of the newest current row (row y + R) in the current ring
  uint32_t phase = 0;    // bit j: parity of the next completion of barrier j
  int jslot = 0;
  const bool partial_bottom = (H & 15) != 0;
#pragma unroll 1
  for (int y = y_start, s = 0; y <= y_end; y++, s++) {
    if ((s & (kBoxRows - 1)) == 0) {
      // a new box: the slot of the box before it is free for everybody -> refill it, then wait for ours
      __syncwarp();
      if (lane == 0) load_box();
      mbar_wait(&bar[jslot], (phase >> jslot) & 1u);
      phase ^= 1u << jslot;
      jslot = jslot + 1 == NB ? 0 : jslot + 1;
    }
    const uint4 rv = *reinterpret_cast<const uint4 *>(rd + 16);
    const uint32_t hl = *reinterpret_cast<const uint32_t *>(rd + 12);
    const uint32_t hr = *reinterpret_cast<const uint32_t *>(rd + 32);
    rd = rd + ROWB == rd0 + L::kRefBytes ? rd0 : rd + ROWB;
    // newest current row: in the mirror when the window below it would wrap around
    const uint8_t *cw = cd0 + (coff < 2 * R * CROWB ? coff + DCB * kBoxRows * CROWB : coff);
    coff = coff + CROWB == (uint32_t)(DCB * kBoxRows * CROWB) ? 0u : coff + CROWB;
    const uint32_t raw[6] = {hl, rv.x, rv.y, rv.z, rv.w, hr};   // bytes x0-4 .. x0+19 of reference row y
    // byte-align the reference row to each of the lane's K horizontal offsets
    uint32_t r4[K][4];
    if (G == 1) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        constexpr int kb = 4 - R;           // + k: byte offset of dx = k - R in raw
        const int b = kb + k, wi = b >> 2;
#pragma unroll
        for (int q = 0; q < 4; q++)
          r4[k][q] = (b & 3) ? __funnelshift_r(raw[wi + q], raw[wi + q + 1 > 5 ? 5 : wi + q + 1], 8u * (uint32_t)(b & 3))
                             : raw[wi + q];
      }
    } else {
      // the lane's first dx starts at byte b = gi*K - R + 4 of raw: align once, then shift by k bytes
      const int b = gi * K - R + 4;
      const bool up = b >= 4;
      const uint32_t sh = 8u * (uint32_t)(b & 3);
      uint32_t base[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const uint32_t lo = up ? raw[j + 1] : raw[j];
        const uint32_t hi = up ? raw[j + 2 > 5 ? 5 : j + 2] : raw[j + 1];
        base[j] = __funnelshift_r(lo, hi, sh);
      }
#pragma unroll
      for (int k = 0; k < K; k++)
#pragma unroll
        for (int q = 0; q < 4; q++) r4[k][q] = k ? __funnelshift_r(base[q], base[q + 1], 8u * (uint32_t)k) : base[q];
    }
#pragma unroll
    for (int k = 0; k < K; k++)
      S[k] += __dp4a(r4[k][0], r4[k][0], __dp4a(r4[k][1], r4[k][1], 0u)) +
              __dp4a(r4[k][2], r4[k][2], __dp4a(r4[k][3], r4[k][3], 0u));
    // current row y - dy = y + R - d against reference row y, for every dy and dx
#pragma unroll
    for (int d = 0; d < ND; d++) {
      const uint4 cv = *reinterpret_cast<const uint4 *>(cw - d * CROWB);
      if (d == 0)   // the row that entered the window: its energy goes into the prefix of sum cur^2
        Sc += __dp4a(cv.x, cv.x, __dp4a(cv.y, cv.y, 0u)) + __dp4a(cv.z, cv.z, __dp4a(cv.w, cv.w, 0u));
#pragma unroll
      for (int k = 0; k < K; k++) {
        uint32_t a = acc[k][d];
        a = __dp4a(cv.x, r4[k][0], a);
        a = __dp4a(cv.y, r4[k][1], a);
        a = __dp4a(cv.z, r4[k][2], a);
        a = __dp4a(cv.w, r4[k][3], a);
        acc[k][d] = a;
      }
    }
    // ---- warp-uniform events of this step (S already includes row y: a candidate that finishes
    // here covers rows up to y, one that starts at the next step begins after it).  At most two
    // candidate rows finish per step: dy = tt - R, whose current row y - dy is row 15 of a block, and
    // -- next to a partial bottom block -- the one whose current row is the frame's last row.
    const int tt = (y + 1 + R) & 15;
    const int ev0 = tt < ND ? tt : -1;
    const int ev1 = (partial_bottom && y >= H - 1 - R) ? y - (H - 1) + R : -1;
    if (ev0 >= 0 || ev1 >= 0) {
#pragma unroll 1
      for (int e = 0; e < 2; e++) {
        const int di = e == 0 ? ev0 : ev1;
        if (di >= 0) {
          const int by = finish(di, y);
          if (by >= 0) publish(by);
        }
      }
    }
    const int yc = y + R;                       // the current row that entered: does it close a block?
    if (yc < H && (((yc & 15) == 15) || yc == H - 1)) {
      const uint32_t a = Sc - ScStart;
      if ((yc >> 4) & 1) A1 = a; else A0 = a;
      ScStart = Sc;
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

template <int R, int K, int G, int NB, int MINB>
cudaError_t launch_stream(const Geom &g, const Frames &f, int npairs, const Out &o, cudaStream_t s) {
  using L = StreamLayout<R, K, G, NB>;
  constexpr int CPW = L::CPW;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  StreamParams sp;
  sp.ncg = (g.nbx + CPW - 1) / CPW;
  // Stripes per pair.  Every warp runs one stripe; stripes are cut evenly (block rows i*n/k).  Fewer stripes
  // re-read fewer rows (2R per stripe) and fill fewer rings, but the warps of a wave finish staggered (the
  // scheduler favours the oldest warp, and one or two warps cannot fill the FMA-heavy pipe), which costs about
  // a third of a stripe's time at the end of the launch -- so several short waves, in which finished CTAs are
  // replaced at once, beat one tall wave (measured, 256 pairs of 1080p +-2: 2 stripes 541 k, 9 stripes 628 k
  // frames/s; 64 pairs: 9 stripes 509 k, 17 stripes 515 k).
  const long long resident = (long long)sms * MINB * kStreamWarps;
  double best_cost = -1.0;
  int best_n = 1;
  for (int n = 1; n <= g.by_count; n++) {
    const int rows = (g.by_count + n - 1) / n;   // block rows of the tallest stripe
    const long long total = (long long)npairs * n * sp.ncg;
    const long long waves = (total + resident - 1) / resident;
    const double cost = ((double)waves + 0.35) * (double)(rows * 16 + 2 * R + 12);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = n; }
  }
  if (const char *e = getenv("ME_B200_STREAM_STRIPES")) {
    const int v = atoi(e);
    if (v >= 1) best_n = v < g.by_count ? v : g.by_count;
  }
  sp.nstripes = best_n;
  sp.total = (long long)npairs * sp.nstripes * sp.ncg;
  CUtensorMap map_ref, map_cur;
  if (!encode_frames_map(&map_ref, f.ref, g.W, g.H, npairs, f.pitch, f.pair_stride, L::ROWB, kBoxRows, true) ||
      !encode_frames_map(&map_cur, f.cur, g.W, g.H, npairs, f.pitch, f.pair_stride, L::CROWB, kBoxRows, true))
    return cudaErrorInvalidValue;
  const int smem = kStreamWarps * L::kWarpBytes;
  auto kern = stream_search_kernel<R, K, G, NB, MINB>;
  static bool attr_set = false;   // per instantiation
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const long long ctas = (sp.total + kStreamWarps - 1) / kStreamWarps;
  kern<<<(unsigned)ctas, kStreamWarps * 32, smem, s>>>(map_ref, map_cur, g, o, sp);
  return cudaGetLastError();
}

bool stream_supported(const Geom &g, const Frames &f, int npairs) {
  if (g.B != 16 || g.R < 1 || g.R > 4 || g.W < 16 || (g.W & 15)) return false;
  if ((f.pitch & 15) || (npairs > 1 && (f.pair_stride & 15))) return false;
  if ((((uintptr_t)f.cur) | ((uintptr_t)f.ref)) & 15) return false;
  if (getenv("ME_B200_NO_STREAM")) return false;
  if (!get_encode()) return false;
  return true;
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
    if (stream_supported(g, ff, n)) {
      // 16x16 blocks, 1 <= R <= 4: the register-streaming kernel (K offsets per lane, G lanes per block column)
      cudaError_t e;
      if (g.R == 1) e = launch_stream<1, 3, 1, 3, 20>(g, ff, n, oo, s);
      else if (g.R == 2) e = launch_stream<2, 5, 1, 3, 16>(g, ff, n, oo, s);
      else if (g.R == 3) e = launch_stream<3, 4, 2, 4, 16>(g, ff, n, oo, s);
      else e = launch_stream<4, 3, 3, 4, 16>(g, ff, n, oo, s);
      if (e != cudaSuccess) return e;
      done += n;
      continue;
    }
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
