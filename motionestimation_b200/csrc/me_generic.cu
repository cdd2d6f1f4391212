// me_generic.cu -- generic exact full-search kernel (any B <= 256, any R, any
// frame size, partial edge blocks, windows larger than shared memory).
//
// One CTA per block.  The clamped window (main.c:69-76) and the current block
// are staged into shared memory with coalesced byte loads when they fit;
// otherwise candidates read the frames through L1/L2.  Each thread walks the
// candidates c = tid, tid + nthreads, ... in the reference's visit order
// (y-major, x-minor, main.c:53-54), computes the exact integer SSD with packed
// byte ops where the row is long enough, forms the reference's float score
// (main.c:27) and keeps the smallest 64-bit key (score bits << 32 | visit index):
// the unsigned minimum of that key IS the reference's "first strict minimum"
// (main.c:56-60).  This kernel is the parity safety net; the tuned kernel
// (me_tiled.cu) handles the benchmark geometries.
#include "me_device.cuh"

namespace me {

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kExactLimit = 1u << 24;  // float holds every integer below this

// Literal restatement of main.c:19-26 for the (adversarial, B > 16 only) case
// where the integer SSD is >= 2^24 and float accumulation starts to round.
__device__ float float_sum_literal(const uint8_t *cur, int cpitch, const uint8_t *cand, int wpitch,
                                   int w, int h) {
  float sum = 0.0f;
  for (int oy = 0; oy < h; oy++)
    for (int ox = 0; ox < w; ox++) {
      int d = (int)cur[oy * cpitch + ox] - (int)cand[oy * wpitch + ox];
      sum = __fadd_rn(sum, (float)(d * d));
    }
  return sum;
}

__device__ __forceinline__ uint32_t ssd_bytes(const uint8_t *cur, int cpitch, const uint8_t *cand,
                                              int wpitch, int w, int h) {
  uint32_t s = 0;
  for (int oy = 0; oy < h; oy++) {
    const uint8_t *a = cur + oy * cpitch;
    const uint8_t *b = cand + oy * wpitch;
    int ox = 0;
    // cur rows are 4-byte aligned in shared memory; candidate rows are not:
    // assemble 4 candidate bytes, then VABSDIFF4 + IDP.4A (4 pixels / 2 int ops).
    for (; ox + 4 <= w; ox += 4) {
      uint32_t av = (uint32_t)a[ox] | ((uint32_t)a[ox + 1] << 8) | ((uint32_t)a[ox + 2] << 16) |
                    ((uint32_t)a[ox + 3] << 24);
      uint32_t bv = (uint32_t)b[ox] | ((uint32_t)b[ox + 1] << 8) | ((uint32_t)b[ox + 2] << 16) |
                    ((uint32_t)b[ox + 3] << 24);
      uint32_t d = __vabsdiffu4(av, bv);
      s = __dp4a(d, d, s);
    }
    for (; ox < w; ox++) {
      int d = (int)a[ox] - (int)b[ox];
      s += (uint32_t)(d * d);
    }
  }
  return s;
}

__global__ void __launch_bounds__(kThreads)
generic_search_kernel(Geom g, Frames f, Out o, int smem_window_ok) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ unsigned long long warp_best[kThreads / 32];
  __shared__ uint32_t warp_ssd[kThreads / 32];

  const int bi_local = blockIdx.x;  // block within the band
  const int pair = blockIdx.y;
  const int bx = bi_local % g.nbx;
  const int by = g.by_begin + bi_local / g.nbx;
  const int bi = by * g.nbx + bx;
  const int x0 = bx * g.B, y0 = by * g.B;
  const int w = min(g.B, g.W - x0), h = min(g.B, g.H - y0);
  // clamped window, inclusive bounds (main.c:73-76)
  const int wx0 = max(0, x0 - g.R), wy0 = max(0, y0 - g.R);
  const int wx1 = min(g.W - 1, x0 + w - 1 + g.R), wy1 = min(g.H - 1, y0 + h - 1 + g.R);
  const int ncx = wx1 - w + 1 - wx0 + 1, ncy = wy1 - h + 1 - wy0 + 1;
  const int ww = wx1 - wx0 + 1, wh = wy1 - wy0 + 1;

  const uint8_t *cur = f.cur + (size_t)pair * f.pair_stride;
  const uint8_t *ref = f.ref + (size_t)pair * f.pair_stride;

  // stage current block (pitch rounded to 4) and, if it fits, the window
  const int cpitch = (w + 3) & ~3;
  uint8_t *s_cur = smem;
  for (int i = threadIdx.x; i < cpitch * h; i += kThreads) {
    int r = i / cpitch, c = i - r * cpitch;
    s_cur[i] = c < w ? cur[(size_t)(y0 + r) * f.pitch + x0 + c] : 0;
  }
  const uint8_t *wbase;
  int wpitch;
  if (smem_window_ok) {
    uint8_t *s_win = smem + ((cpitch * h + 15) & ~15);
    for (int i = threadIdx.x; i < ww * wh; i += kThreads) {
      int r = i / ww, c = i - r * ww;
      s_win[i] = ref[(size_t)(wy0 + r) * f.pitch + wx0 + c];
    }
    wbase = s_win;
    wpitch = ww;
  } else {
    wbase = ref + (size_t)wy0 * f.pitch + wx0;
    wpitch = (int)f.pitch;
  }
  __syncthreads();

  const float area = (float)(w * h);  // main.c:27: int product converted to float
  unsigned long long best = ~0ull;
  uint32_t best_ssd = 0;
  const int ncand = ncx * ncy;
  for (int c = threadIdx.x; c < ncand; c += kThreads) {
    const int cy = c / ncx, cx = c - cy * ncx;
    const uint8_t *cand = wbase + cy * wpitch + cx;
    const uint32_t s = ssd_bytes(s_cur, cpitch, cand, wpitch, w, h);
    const float sum = s < kExactLimit ? (float)s : float_sum_literal(s_cur, cpitch, cand, wpitch, w, h);
    const float sc = __fdiv_rn(sum, area);
    const unsigned long long key = ((unsigned long long)score_bits(sc) << 32) | (uint32_t)c;
    if (key < best) {
      best = key;
      best_ssd = s;
    }
  }
  // CTA arg-min of the key; the ssd rides along
  for (int off = 16; off; off >>= 1) {
    unsigned long long ok = __shfl_down_sync(0xffffffffu, best, off);
    uint32_t os = __shfl_down_sync(0xffffffffu, best_ssd, off);
    if (ok < best) {
      best = ok;
      best_ssd = os;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    warp_best[threadIdx.x >> 5] = best;
    warp_ssd[threadIdx.x >> 5] = best_ssd;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kThreads / 32; i++)
      if (warp_best[i] < best) {
        best = warp_best[i];
        best_ssd = warp_ssd[i];
      }
    const int c = (int)(uint32_t)best;
    const int cy = c / ncx, cx = c - cy * ncx;
    const size_t oi = (size_t)pair * g.nbx * g.nby + bi;
    if (o.mvx) o.mvx[oi] = wx0 + cx - x0;  // main.c:58
    if (o.mvy) o.mvy[oi] = wy0 + cy - y0;  // main.c:59
    if (o.ssd) o.ssd[oi] = best_ssd;
    if (o.score) o.score[oi] = __uint_as_float((uint32_t)(best >> 32));
  }
}

}  // namespace

cudaError_t launch_generic(const Geom &g, const Frames &f, int npairs, const Out &o, cudaStream_t s) {
  const int cpitch = (g.B + 3) & ~3;
  const size_t cur_bytes = ((size_t)cpitch * g.B + 15) & ~(size_t)15;
  const size_t win_side = (size_t)g.B + 2 * (size_t)g.R;
  const size_t win_bytes = win_side * win_side;
  static const size_t kMaxSmem = 200 * 1024;
  const int ok = cur_bytes + win_bytes <= kMaxSmem;
  const size_t smem = ok ? cur_bytes + win_bytes : cur_bytes;
  cudaError_t e = cudaFuncSetAttribute(generic_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kMaxSmem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)(g.nbx * g.by_count), (unsigned)npairs);
  generic_search_kernel<<<grid, kThreads, smem, s>>>(g, f, o, ok);
  return cudaGetLastError();
}

}  // namespace me
