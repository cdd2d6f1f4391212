// me_fast.cu -- three-step and diamond search as warp-cooperative kernels for sm_100a
// (BASELINE.json config 4).
//
// PARITY UNPINNED: the reference has no fast search (only the exhaustive scans of
// src/cpu/main.c and src/cpu/main_ssim.c), so there is nothing of the reference's to be
// bit-exact with.  The two patterns are therefore DEFINED here (and restated on the CPU by the
// repo's checker, which the GPU tests compare against bit for bit), reusing every rule the
// reference does have:
//   grid, partial edge blocks   src/common/prediction_frame.c:9-23
//   clamped search window       src/cpu/main.c:69-76   (a point outside it does not exist)
//   cost                        src/cpu/main.c:18-27   float(sum (cur-ref)^2) / float(w*h)
//   comparison                  src/cpu/main.c:56      strict '<': an earlier point keeps a tie
//   mv                          src/cpu/main.c:58-59
// Three-step (Koga et al. 1981): the zero-motion candidate is the incumbent; step S = largest
//   power of two <= max(1, (R+1)/2); the 8 neighbours of the centre at distance S are visited in
//   raster order (dy = -S, 0, +S outer; dx inner) and replace the incumbent on a strictly smaller
//   score; the best point becomes the centre, S halves, until S = 0.
// Diamond (Zhu & Ma 2000): large diamond = (0,+-2) (+-2,0) (+-1,+-1) around the incumbent centre,
//   visited in raster order; while a neighbour wins, the centre moves there and the large diamond
//   repeats; once the centre keeps the minimum, the small diamond (0,+-1) (+-1,0) is evaluated
//   once and its best point is the result.
//
// One warp per block.  The block's pixels are staged once in shared memory; every step
// evaluates the (up to) 8 points of the pattern around the incumbent centre in parallel:
// 4 lanes per point, each lane takes every fourth row, VABSDIFF4.U8 + IDP.4A.U8.U8 on words
// assembled from two aligned loads and a funnel shift (byte loads for odd widths / unaligned
// layouts).  The reference rows come straight from global memory: which rows are needed is
// data dependent, the frames stay L2 resident, and a step is latency bound, not bandwidth bound.
// The 8 scores + the incumbent are combined with one 64-bit warp minimum of
// (score bits << 8 | visit rank), the incumbent having rank 0: exactly "first strictly
// smaller in visit order".
#include <cuda_runtime.h>
#include <stdint.h>

#include "me_device.cuh"

namespace me {

namespace {

constexpr int kFastWarps = 8;

__constant__ int kSquare8[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
__constant__ int kLarge8[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {2, 0}, {-1, 1}, {1, 1}, {0, 2}};
__constant__ int kSmall4[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};

struct BlockCtx {
  const uint8_t *ref;   // reference frame of this pair
  const uint8_t *s_cur; // staged current block, row pitch cp (multiple of 4)
  size_t pitch;
  int cp, w, h;
  int x0, y0;
  int lo_x, hi_x, lo_y, hi_y;  // inclusive bounds of a candidate's top-left corner
  float area;
  bool words;           // aligned layout and w % 4 == 0: packed-byte path
};

// partial SSD of the rows r = sub, sub + 4, ... of the candidate at (x, y)
__device__ __forceinline__ uint32_t partial_ssd(const BlockCtx &c, int x, int y, int sub) {
  uint32_t s = 0;
  if (c.words) {
    const uint32_t shift = 8u * (uint32_t)(x & 3);
    for (int r = sub; r < c.h; r += 4) {
      const uint32_t *rp = reinterpret_cast<const uint32_t *>(c.ref + (size_t)(y + r) * c.pitch + (x & ~3));
      const uint32_t *cp = reinterpret_cast<const uint32_t *>(c.s_cur + r * c.cp);
      uint32_t lo = __ldg(rp);
      for (int k = 0; k < (c.w >> 2); k++) {
        // the word after the last one is only touched when the row really extends into it
        const uint32_t hi = (shift != 0u || k + 1 < (c.w >> 2)) ? __ldg(rp + k + 1) : 0u;
        const uint32_t d = __vabsdiffu4(cp[k], __funnelshift_r(lo, hi, shift));
        s = __dp4a(d, d, s);
        lo = hi;
      }
    }
  } else {
    for (int r = sub; r < c.h; r += 4) {
      const uint8_t *rp = c.ref + (size_t)(y + r) * c.pitch + x;
      const uint8_t *cp = c.s_cur + r * c.cp;
      for (int k = 0; k < c.w; k++) {
        const int d = (int)cp[k] - (int)__ldg(rp + k);
        s += (uint32_t)(d * d);
      }
    }
  }
  return s;
}

// main.c:19-27 literally (float accumulation in raster order): only needed once SSD >= 2^24
__device__ float literal_score(const BlockCtx &c, int x, int y) {
  float sum = 0.0f;
  for (int r = 0; r < c.h; r++)
    for (int k = 0; k < c.w; k++) {
      const int d = (int)c.s_cur[r * c.cp + k] - (int)__ldg(c.ref + (size_t)(y + r) * c.pitch + x + k);
      sum = __fadd_rn(sum, (float)(d * d));
    }
  return __fdiv_rn(sum, c.area);
}

struct Best {
  float score;
  uint32_t ssd;
  int x, y;
};

// Evaluate npts pattern points around (cx, cy) (offsets scaled by `scale`), 4 lanes per point,
// and fold them into the incumbent.  Returns the number of points that were evaluated.
__device__ __forceinline__ int visit(const BlockCtx &c, const int (*off)[2], int npts, int scale, int cx, int cy,
                                     Best &b, int lane) {
  const int pt = lane >> 2, sub = lane & 3;
  int x = cx, y = cy;
  bool ok = false;
  if (pt < npts) {
    x = cx + off[pt][0] * scale;
    y = cy + off[pt][1] * scale;
    ok = x >= c.lo_x && x <= c.hi_x && y >= c.lo_y && y <= c.hi_y;
  }
  uint32_t s = ok ? partial_ssd(c, x, y, sub) : 0u;
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  float sc = 0.0f;
  if (ok) sc = s < (1u << 24) ? __fdiv_rn((float)s, c.area) : literal_score(c, x, y);
  // rank 0 = the incumbent (keeps ties), points follow in visit order
  unsigned long long key = ok ? (((unsigned long long)__float_as_uint(sc) << 8) | (unsigned)(pt + 1)) : ~0ull;
  if (lane == 0) {
    const unsigned long long inc = (unsigned long long)__float_as_uint(b.score) << 8;
    // lane 0 carries point 0 as well: keep whichever is smaller (the incumbent on a tie)
    key = inc <= key ? inc : key;
  }
  unsigned long long m = key;
  for (int o = 16; o; o >>= 1) {
    const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
    m = v < m ? v : m;
  }
  const int win = (int)(m & 0xffu);  // 0: the incumbent stays
  if (win != 0) {
    const int src = (win - 1) * 4;
    b.score = __uint_as_float((uint32_t)(m >> 8));
    b.ssd = __shfl_sync(0xffffffffu, s, src);
    b.x = __shfl_sync(0xffffffffu, x, src);
    b.y = __shfl_sync(0xffffffffu, y, src);
  }
  return __popc(__ballot_sync(0xffffffffu, ok && sub == 0));
}

__global__ void __launch_bounds__(kFastWarps * 32)
fast_search_kernel(Geom g, Frames f, Out o, int algo, int first_step, int nblocks_band,
                   unsigned long long *evals) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bl = blockIdx.x * kFastWarps + warp;  // block within the band
  const int pair = blockIdx.y;
  if (bl >= nblocks_band) return;
  const int bx = bl % g.nbx, by = g.by_begin + bl / g.nbx;

  BlockCtx c;
  c.x0 = bx * g.B;
  c.y0 = by * g.B;
  c.w = min(g.B, g.W - c.x0);
  c.h = min(g.B, g.H - c.y0);
  c.cp = (g.B + 3) & ~3;
  c.pitch = f.pitch;
  c.ref = f.ref + (size_t)pair * f.pair_stride;
  const uint8_t *cur = f.cur + (size_t)pair * f.pair_stride;
  uint8_t *sc = smem + (size_t)warp * c.cp * g.B;
  for (int i = lane; i < c.cp * c.h; i += 32) {
    const int r = i / c.cp, k = i - r * c.cp;
    sc[i] = k < c.w ? cur[(size_t)(c.y0 + r) * f.pitch + c.x0 + k] : 0;
  }
  __syncwarp();
  c.s_cur = sc;
  // clamped window (main.c:73-76) as bounds of the candidate's top-left corner (main.c:53-54)
  c.lo_x = max(0, c.x0 - g.R);
  c.lo_y = max(0, c.y0 - g.R);
  c.hi_x = min(g.W - 1, c.x0 + c.w - 1 + g.R) - c.w + 1;
  c.hi_y = min(g.H - 1, c.y0 + c.h - 1 + g.R) - c.h + 1;
  c.area = (float)(c.w * c.h);
  c.words = (c.w & 3) == 0 && (f.pitch & 3) == 0 && (f.pair_stride & 3) == 0 && (((uintptr_t)f.ref) & 3) == 0;

  // incumbent: zero motion (always inside the window)
  Best b;
  b.x = c.x0;
  b.y = c.y0;
  {
    uint32_t s = (lane < 4) ? partial_ssd(c, b.x, b.y, lane) : 0u;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s = __shfl_sync(0xffffffffu, s, 0);
    b.ssd = s;
    b.score = s < (1u << 24) ? __fdiv_rn((float)s, c.area) : literal_score(c, b.x, b.y);
  }
  int nev = 1;
  if (algo == 1) {
    for (int st = first_step; st >= 1; st >>= 1) nev += visit(c, kSquare8, 8, st, b.x, b.y, b, lane);
  } else {
    for (;;) {
      const int cx = b.x, cy = b.y;
      nev += visit(c, kLarge8, 8, 1, cx, cy, b, lane);
      if (b.x == cx && b.y == cy) break;
    }
    nev += visit(c, kSmall4, 4, 1, b.x, b.y, b, lane);
  }
  if (lane == 0) {
    const size_t oi = (size_t)pair * g.nbx * g.nby + (size_t)by * g.nbx + bx;
    if (o.mvx) o.mvx[oi] = b.x - c.x0;   // main.c:58
    if (o.mvy) o.mvy[oi] = b.y - c.y0;   // main.c:59
    if (o.ssd) o.ssd[oi] = b.ssd;
    if (o.score) o.score[oi] = b.score;
    if (evals) atomicAdd(evals, (unsigned long long)nev);
  }
}

}  // namespace

int tss_first_step(int R) {
  int half = (R + 1) / 2, s = 1;
  if (half < 1) half = 1;
  while (s * 2 <= half) s *= 2;
  return R > 0 ? s : 0;
}

// algo: 1 = three-step, 2 = diamond.  evals (device, may be null) accumulates the number of
// candidate evaluations.
cudaError_t launch_fast(const Geom &g, const Frames &f, int npairs, const Out &o, int algo,
                        unsigned long long *evals, cudaStream_t s) {
  const int nblocks_band = g.nbx * g.by_count;
  if (nblocks_band <= 0 || npairs <= 0) return cudaSuccess;
  const int cp = (g.B + 3) & ~3;
  const size_t smem = (size_t)kFastWarps * cp * g.B;
  cudaError_t e = cudaFuncSetAttribute(fast_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  for (int done = 0; done < npairs; done += 65535) {
    const int np = npairs - done > 65535 ? 65535 : npairs - done;
    Frames ff = f;
    ff.cur += (size_t)done * f.pair_stride;
    ff.ref += (size_t)done * f.pair_stride;
    Out oo = o;
    const size_t off = (size_t)done * g.nbx * g.nby;
    if (oo.mvx) oo.mvx += off;
    if (oo.mvy) oo.mvy += off;
    if (oo.ssd) oo.ssd += off;
    if (oo.score) oo.score += off;
    dim3 grid((unsigned)((nblocks_band + kFastWarps - 1) / kFastWarps), (unsigned)np);
    fast_search_kernel<<<grid, kFastWarps * 32, smem, s>>>(g, ff, oo, algo, tss_first_step(g.R), nblocks_band, evals);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace me
