/*
 * me_oracle.c -- CPU restatement of the reference full-search MSE path.
 * TEST INFRASTRUCTURE ONLY (see me_oracle.h).  Parity status: PINNED.
 *
 * The arithmetic follows the reference line by line:
 *   cost      src/cpu/main.c:18-36   float sum += (int diff)*(int diff); score = sum / (w*h)
 *   scan      src/cpu/main.c:53-62   y outer, x inner, strict '<' against INFINITY-initialised best
 *   window    src/cpu/main.c:69-76   clamp [tl-R, br+R] to the frame
 *   mv        src/cpu/main.c:58-59   x - top_left_x, y - top_left_y
 *   grid      src/common/prediction_frame.c:9-23
 *
 * One shortcut, proven equivalent and cross-checked by tests against the
 * literal loop (env ME_ORACLE_LITERAL=1 forces the literal loop everywhere):
 * the reference accumulates integer squares into a float.  While the running
 * sum stays below 2^24 every partial sum is an integer a float holds exactly,
 * so the float sum equals the integer SSD; once the exact sum reaches 2^24
 * the float sum is >= 2^24 too (round-to-nearest is monotone and 2^24 is
 * representable).  Hence: integer SSD < 2^24  ==> float sum == (float)SSD,
 * and only candidates with SSD >= 2^24 need the literal float loop.
 */
#include "me_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static int g_literal = -1;
static int literal_mode(void) {
  if (g_literal < 0) {
    const char *e = getenv("ME_ORACLE_LITERAL");
    g_literal = (e && e[0] == '1') ? 1 : 0;
  }
  return g_literal;
}

int me_oracle_num_blocks(int width, int height, int blk_dim) {
  if (width <= 0 || height <= 0 || blk_dim <= 0) return 0;
  int nbx = (width + blk_dim - 1) / blk_dim;   /* prediction_frame.c:9 */
  int nby = (height + blk_dim - 1) / blk_dim;  /* prediction_frame.c:10 */
  return nbx * nby;
}

void me_oracle_block_geom(int i, int width, int height, int blk_dim,
                          int *x0, int *y0, int *w, int *h) {
  int nbx = (width + blk_dim - 1) / blk_dim;
  int bx = i % nbx, by = i / nbx;              /* prediction_frame.c:15-16 */
  *x0 = bx * blk_dim;
  *y0 = by * blk_dim;
  *w = (*x0 + blk_dim) < width ? blk_dim : width - *x0;    /* :20 */
  *h = (*y0 + blk_dim) < height ? blk_dim : height - *y0;  /* :21 */
}

/* main.c:18-27, literal: float accumulation of int products in raster order. */
static float sum_literal(const uint8_t *cur, const uint8_t *ref, int stride,
                         int bx0, int by0, int cx, int cy, int w, int h) {
  float sum = 0;
  for (int oy = 0; oy < h; oy++)
    for (int ox = 0; ox < w; ox++) {
      int a = cur[(by0 + oy) * stride + bx0 + ox];
      int b = ref[(cy + oy) * stride + cx + ox];
      sum += (a - b) * (a - b);
    }
  return sum;
}

static uint32_t ssd_int(const uint8_t *cur, const uint8_t *ref, int stride,
                        int bx0, int by0, int cx, int cy, int w, int h) {
  uint32_t s = 0; /* w*h*255^2 < 2^32 for w*h <= 66051; larger blocks use 64-bit below */
  for (int oy = 0; oy < h; oy++) {
    const uint8_t *pa = cur + (by0 + oy) * stride + bx0;
    const uint8_t *pb = ref + (cy + oy) * stride + cx;
    uint32_t rs = 0;
    for (int ox = 0; ox < w; ox++) {
      int d = (int)pa[ox] - (int)pb[ox];
      rs += (uint32_t)(d * d);
    }
    s += rs;
  }
  return s;
}

static void search_block(const uint8_t *cur, const uint8_t *ref, int W, int H,
                         int B, int R, int i, me_oracle_result *out) {
  int x0, y0, w, h;
  me_oracle_block_geom(i, W, H, B, &x0, &y0, &w, &h);
  int brx = x0 + w - 1, bry = y0 + h - 1;                   /* block.c:10-11 */
  int wx0 = (x0 - R) < 0 ? 0 : x0 - R;                      /* main.c:73 */
  int wy0 = (y0 - R) < 0 ? 0 : y0 - R;                      /* main.c:74 */
  int wx1 = (brx + R) >= W ? W - 1 : brx + R;               /* main.c:75 */
  int wy1 = (bry + R) >= H ? H - 1 : bry + R;               /* main.c:76 */
  int lit = literal_mode() || ((uint64_t)w * (uint64_t)h > 66051u);
  float best = INFINITY;                                    /* main.c:43 */
  float bdx = 0, bdy = 0;
  uint32_t bssd = 0;
  float area = (float)(w * h);                              /* main.c:27 int -> float */
  for (int y = wy0; y <= wy1 - h + 1; y++) {                /* main.c:53 */
    for (int x = wx0; x <= wx1 - w + 1; x++) {              /* main.c:54 */
      float sum;
      uint32_t s = 0;
      if (lit) {
        sum = sum_literal(cur, ref, W, x0, y0, x, y, w, h);
      } else {
        s = ssd_int(cur, ref, W, x0, y0, x, y, w, h);
        sum = (s < (1u << 24)) ? (float)s : sum_literal(cur, ref, W, x0, y0, x, y, w, h);
      }
      float score = sum / area;                             /* main.c:27 */
      if (score < best) {                                   /* main.c:56 */
        best = score;
        bdx = (float)(x - x0);                              /* main.c:58 */
        bdy = (float)(y - y0);                              /* main.c:59 */
        bssd = lit ? 0xffffffffu : s;
      }
    }
  }
  out->mvx = (int)bdx;                                      /* main.c:79 */
  out->mvy = (int)bdy;
  out->score = best;
  if (bssd == 0xffffffffu && lit) {
    /* recompute the winner's exact integer SSD (saturating for huge blocks) */
    uint64_t s64 = 0;
    for (int oy = 0; oy < h; oy++)
      for (int ox = 0; ox < w; ox++) {
        int d = (int)cur[(y0 + oy) * W + x0 + ox] -
                (int)ref[(y0 + out->mvy + oy) * W + x0 + out->mvx + ox];
        s64 += (uint64_t)(d * d);
      }
    bssd = s64 > 0xffffffffull ? 0xffffffffu : (uint32_t)s64;
  }
  out->ssd = bssd;
}

typedef struct job {
  const uint8_t *cur, *ref;
  int W, H, B, R, begin, end, base;
  me_oracle_result *out;
} job;

static void *job_main(void *p) {
  job *j = (job *)p;
  for (int i = j->begin; i < j->end; i++)
    search_block(j->cur, j->ref, j->W, j->H, j->B, j->R, i, &j->out[i - j->base]);
  return NULL;
}

int me_oracle_search(const uint8_t *cur, const uint8_t *ref,
                     int width, int height, int blk_dim, int extra_span,
                     int blk_begin, int blk_end, int nthreads,
                     me_oracle_result *out) {
  int nb = me_oracle_num_blocks(width, height, blk_dim);
  if (!cur || !ref || !out || nb <= 0 || extra_span < 0) return -1;
  if (blk_begin < 0 || blk_end > nb || blk_begin > blk_end) return -1;
  int n = blk_end - blk_begin;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  if (nthreads == 1) {
    job j = {cur, ref, width, height, blk_dim, extra_span, blk_begin, blk_end, blk_begin, out};
    job_main(&j);
    return 0;
  }
  /* interleave small chunks so border (cheap) and interior (dear) blocks mix */
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  job *jobs = (job *)malloc(sizeof(job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    int b = blk_begin + (int)((int64_t)n * t / nthreads);
    int e = blk_begin + (int)((int64_t)n * (t + 1) / nthreads);
    job j = {cur, ref, width, height, blk_dim, extra_span, b, e, blk_begin, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}

/* candidates along one axis for a block at p with extent e in a frame of N
 * (main.c:53-54 with the clamps of :73-76) */
static int axis_cands(int p, int e, int N, int R) {
  int lo = p - R < 0 ? 0 : p - R;
  int hi = (p + e - 1 + R) >= N ? N - 1 : p + e - 1 + R;
  return hi - e + 1 - lo + 1;
}

uint64_t me_oracle_candidates(int W, int H, int B, int R) {
  uint64_t sx = 0, sy = 0;
  for (int x0 = 0; x0 < W; x0 += B) {
    int w = x0 + B < W ? B : W - x0;
    sx += (uint64_t)axis_cands(x0, w, W, R);
  }
  for (int y0 = 0; y0 < H; y0 += B) {
    int h = y0 + B < H ? B : H - y0;
    sy += (uint64_t)axis_cands(y0, h, H, R);
  }
  return sx * sy;
}

uint64_t me_oracle_pixel_compares(int W, int H, int B, int R) {
  uint64_t sx = 0, sy = 0;
  for (int x0 = 0; x0 < W; x0 += B) {
    int w = x0 + B < W ? B : W - x0;
    sx += (uint64_t)w * (uint64_t)axis_cands(x0, w, W, R);
  }
  for (int y0 = 0; y0 < H; y0 += B) {
    int h = y0 + B < H ? B : H - y0;
    sy += (uint64_t)h * (uint64_t)axis_cands(y0, h, H, R);
  }
  return sx * sy;
}

int me_oracle_motion_compensate(const uint8_t *ref, int W, int H, int B,
                                const me_oracle_result *res, uint8_t *mc) {
  int nb = me_oracle_num_blocks(W, H, B);
  int bad = 0;
  for (int i = 0; i < nb; i++) {
    int x0, y0, w, h;
    me_oracle_block_geom(i, W, H, B, &x0, &y0, &w, &h);
    int cx0 = x0 + res[i].mvx, cy0 = y0 + res[i].mvy;      /* utils.c:112-113 */
    for (int ox = 0; ox < w; ox++)
      for (int oy = 0; oy < h; oy++) {
        int cx = cx0 + ox, cy = cy0 + oy;
        if (cx >= 0 && cy >= 0 && cx < W && cy < H)          /* utils.c:122 */
          mc[(y0 + oy) * W + x0 + ox] = ref[cy * W + cx];
        else
          bad = 1;
      }
  }
  return bad ? -1 : 0;
}

void me_oracle_frame_diff(const uint8_t *a, const uint8_t *b, int n, uint8_t *out) {
  for (int i = 0; i < n; i++) {
    int d = (int)a[i] - (int)b[i];                           /* utils.c:96-98 */
    out[i] = (uint8_t)(d < 0 ? -d : d);
  }
}

double me_oracle_psnr(const uint8_t *a, const uint8_t *b, int W, int H) {
  double mse = 0.0;
  int mx = 0;
  for (int i = 0; i < W * H; i++) {                           /* utils.c:146-156 */
    if (mx < a[i]) mx = a[i];
    if (mx < b[i]) mx = b[i];
    double t = abs((int)a[i] - (int)b[i]);
    mse += t * t;
  }
  mse /= W * H;                                               /* utils.c:157 */
  if (mse == 0) return 99.0;                                  /* utils.c:160 */
  return 20 * log10(mx) - 10 * log10(mse);                    /* utils.c:162 */
}
