/*
 * me_oracle_fast.c -- CPU definition of the two fast block-matching searches.
 * TEST INFRASTRUCTURE ONLY (see me_oracle.h).
 *
 * Parity status: PARITY UNPINNED.  The reference contains no fast search at all (only the
 * exhaustive scans of src/cpu/main.c and src/cpu/main_ssim.c), so
 * there is no reference output, golden vector or test to pin these against.  BASELINE.json
 * config 4 names "diamond / three-step", hence this file DEFINES them, reusing every rule the
 * reference does have:
 *   grid, partial edge blocks   src/common/prediction_frame.c:9-23
 *   clamped search window       src/cpu/main.c:69-76   (a point outside it does not exist)
 *   cost                        src/cpu/main.c:18-27   float(sum (cur-ref)^2) / float(w*h)
 *   comparison                  src/cpu/main.c:56      strict '<': an earlier point keeps a tie
 *   mv                          src/cpu/main.c:58-59
 * and the textbook patterns:
 *   three-step search (Koga et al. 1981): step S = largest power of two <= max(1, (R+1)/2);
 *     the centre is the incumbent, its 8 neighbours at distance S are visited in raster order
 *     (dy = -S, 0, +S outer; dx = -S, 0, +S inner) and replace the incumbent on strictly
 *     smaller score; the best point becomes the centre, S halves, until S = 0.
 *   diamond search (Zhu & Ma 2000): large diamond = centre + (0,+-2) (+-2,0) (+-1,+-1), visited
 *     in raster order after the incumbent centre; while a neighbour wins the centre moves there
 *     and the large diamond repeats; when the centre keeps the minimum the small diamond
 *     (0,+-1) (+-1,0) is evaluated once and the best of it is the result.
 * Both never leave the clamped window [x0-R, x0+R] x [y0-R, y0+R] intersected with the frame.
 * The CUDA kernels (motionestimation_b200/csrc/me_fast.cu) must reproduce this file bit for bit.
 */
#include "me_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

typedef struct fast_ctx {
  const uint8_t *cur, *ref;
  int W, H, x0, y0, w, h;
  int lo_x, hi_x, lo_y, hi_y; /* inclusive bounds of the candidate's top-left corner */
  float area;
  uint64_t evals;
} fast_ctx;

/* score + exact SSD of the candidate at (x, y); the float sum is exact below 2^24 (me_oracle.c) */
static float point_score(fast_ctx *c, int x, int y, uint32_t *ssd_out) {
  uint64_t s = 0;
  float fsum = 0;
  for (int oy = 0; oy < c->h; oy++)
    for (int ox = 0; ox < c->w; ox++) {
      const int d = (int)c->cur[(c->y0 + oy) * c->W + c->x0 + ox] - (int)c->ref[(y + oy) * c->W + x + ox];
      s += (uint64_t)(d * d);
      fsum += d * d;                                         /* main.c:24, literal */
    }
  c->evals++;
  *ssd_out = s > 0xffffffffull ? 0xffffffffu : (uint32_t)s;
  return fsum / c->area;                                     /* main.c:27 */
}

static int inside(const fast_ctx *c, int x, int y) {
  return x >= c->lo_x && x <= c->hi_x && y >= c->lo_y && y <= c->hi_y;
}

/* visit `n` offsets around (cx, cy) in the given (raster) order; strict '<' against the incumbent */
static void visit(fast_ctx *c, int cx, int cy, const int (*off)[2], int n, int scale,
                  float *best, uint32_t *best_ssd, int *bx, int *by) {
  for (int k = 0; k < n; k++) {
    const int x = cx + off[k][0] * scale, y = cy + off[k][1] * scale;
    if (!inside(c, x, y)) continue;
    uint32_t ssd;
    const float s = point_score(c, x, y, &ssd);
    if (s < *best) {                                         /* main.c:56 */
      *best = s;
      *best_ssd = ssd;
      *bx = x;
      *by = y;
    }
  }
}

static const int kSquare8[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
static const int kLarge8[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {2, 0}, {-1, 1}, {1, 1}, {0, 2}};
static const int kSmall4[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};

int me_oracle_tss_first_step(int R) {
  int half = (R + 1) / 2, s = 1;
  if (half < 1) half = 1;
  while (s * 2 <= half) s *= 2;
  return R > 0 ? s : 0;
}

static void fast_block(const uint8_t *cur, const uint8_t *ref, int W, int H, int B, int R, int algo, int i,
                       me_oracle_result *out, uint64_t *evals) {
  fast_ctx c;
  c.cur = cur; c.ref = ref; c.W = W; c.H = H; c.evals = 0;
  me_oracle_block_geom(i, W, H, B, &c.x0, &c.y0, &c.w, &c.h);
  const int brx = c.x0 + c.w - 1, bry = c.y0 + c.h - 1;
  const int wx0 = (c.x0 - R) < 0 ? 0 : c.x0 - R;             /* main.c:73 */
  const int wy0 = (c.y0 - R) < 0 ? 0 : c.y0 - R;             /* main.c:74 */
  const int wx1 = (brx + R) >= W ? W - 1 : brx + R;          /* main.c:75 */
  const int wy1 = (bry + R) >= H ? H - 1 : bry + R;          /* main.c:76 */
  c.lo_x = wx0; c.hi_x = wx1 - c.w + 1;                      /* main.c:54 */
  c.lo_y = wy0; c.hi_y = wy1 - c.h + 1;                      /* main.c:53 */
  c.area = (float)(c.w * c.h);
  int bx = c.x0, by = c.y0;                                  /* zero motion is always inside */
  uint32_t bssd;
  float best = point_score(&c, bx, by, &bssd);
  if (algo == 1) {
    for (int s = me_oracle_tss_first_step(R); s >= 1; s >>= 1) {
      const int cx = bx, cy = by;
      visit(&c, cx, cy, kSquare8, 8, s, &best, &bssd, &bx, &by);
    }
  } else {
    for (;;) {
      const int cx = bx, cy = by;
      visit(&c, cx, cy, kLarge8, 8, 1, &best, &bssd, &bx, &by);
      if (bx == cx && by == cy) break;                       /* the centre kept the minimum */
    }
    const int cx = bx, cy = by;
    visit(&c, cx, cy, kSmall4, 4, 1, &best, &bssd, &bx, &by);
  }
  out->mvx = bx - c.x0;                                      /* main.c:58 */
  out->mvy = by - c.y0;                                      /* main.c:59 */
  out->ssd = bssd;
  out->score = best;
  if (evals) *evals += c.evals;
}

typedef struct fast_job {
  const uint8_t *cur, *ref;
  int W, H, B, R, algo, begin, end, base;
  me_oracle_result *out;
  uint64_t evals;
} fast_job;

static void *fast_job_main(void *p) {
  fast_job *j = (fast_job *)p;
  for (int i = j->begin; i < j->end; i++)
    fast_block(j->cur, j->ref, j->W, j->H, j->B, j->R, j->algo, i, &j->out[i - j->base], &j->evals);
  return NULL;
}

/* algo: 1 = three-step, 2 = diamond.  evaluated (may be NULL) receives the number of
 * candidate evaluations (the unit of work of a fast search). */
int me_oracle_search_fast(const uint8_t *cur, const uint8_t *ref, int width, int height, int blk_dim,
                          int extra_span, int algo, int blk_begin, int blk_end, int nthreads,
                          me_oracle_result *out, uint64_t *evaluated) {
  const int nb = me_oracle_num_blocks(width, height, blk_dim);
  if (!cur || !ref || !out || nb <= 0 || extra_span < 0 || (algo != 1 && algo != 2)) return -1;
  if (blk_begin < 0 || blk_end > nb || blk_begin > blk_end) return -1;
  const int n = blk_end - blk_begin;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > n) nthreads = n > 0 ? n : 1;
  pthread_t th[256];
  fast_job jobs[256];
  for (int t = 0; t < nthreads; t++) {
    fast_job j = {cur, ref, width, height, blk_dim, extra_span, algo,
                  blk_begin + (int)((int64_t)n * t / nthreads), blk_begin + (int)((int64_t)n * (t + 1) / nthreads),
                  blk_begin, out, 0};
    jobs[t] = j;
    if (nthreads == 1) fast_job_main(&jobs[0]);
    else pthread_create(&th[t], NULL, fast_job_main, &jobs[t]);
  }
  uint64_t ev = 0;
  for (int t = 0; t < nthreads; t++) {
    if (nthreads > 1) pthread_join(th[t], NULL);
    ev += jobs[t].evals;
  }
  if (evaluated) *evaluated = ev;
  return 0;
}
