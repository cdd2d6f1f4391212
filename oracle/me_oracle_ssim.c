/*
 * me_oracle_ssim.c -- CPU restatement of the reference SSIM-cost full search.
 * TEST INFRASTRUCTURE ONLY (see me_oracle.h).  Parity status: PINNED -- checked block by
 * block (motion vector + score bits) against the UNMODIFIED reference compiled into
 * oracle/_ref/libme_ref_ssim.so (oracle/ref_harness_ssim.c) and against fixtures that the
 * unmodified reference produced (tests/golden/make_golden.py).
 *
 * What is restated (paths relative to the reference checkout):
 *   mean        src/common/ssim.c:3-14    float sum of the pixels / (w*h)
 *   variance    src/common/ssim.c:16-27   float sum of (p - mean)^2 in raster order / (w*h)
 *   cross term  src/common/ssim.c:29-41   the two means arrive as `int` (truncated, ssim.h:12);
 *                                         int products accumulated in a float / (w*h)
 *   score       src/common/ssim.c:44-60   luminance * contrast * structure, all in float,
 *                                         sqrt through double, C1=0.01 C2=0.09 C3=0.045
 *   scan        src/common/ssim.c:83-107  y outer, x inner, strict '>' against a best of 0
 *   window/mv   src/cpu/main_ssim.c:16-30 clamp [tl-R, br+R] to the frame; mv = x-x0, y-y0
 *
 * Float arithmetic is order dependent, so everything that rounds is done literally (this file
 * is compiled with -ffp-contract=off; the reference's build has no FMA either).  Two sums are
 * exact integers and therefore computed as such (ME_ORACLE_LITERAL=1 forces the literal float
 * loops instead; tests cross-check both):
 *   - the pixel sum of the mean: every partial sum is an integer < 2^24 while w*h <= 65793;
 *   - the cross sum: |partial sums| <= w*h*255^2 < 2^24 while w*h <= 258.
 * The statistics of the reference block depend only on the candidate POSITION, not on the
 * block that looks at it; for full-size blocks they are tabulated once per frame (the
 * reference recomputes them for every (block, candidate)), which is what makes 1080p-size
 * checks affordable.
 *
 * Defect kept visible: when no candidate scores above 0 the reference leaves the motion vector
 * uninitialised (ssim.c:88-103, main_ssim.c:26-27).  This restatement -- like the harness, which
 * zero-fills malloc -- reports MV (0,0), score 0 and found = 0 for such blocks.
 */
#include "me_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static int ssim_literal(void) {
  const char *e = getenv("ME_ORACLE_LITERAL");
  return e && e[0] == '1';
}

typedef struct pos_stat {
  float mean;   /* ssim.c:13  */
  float stddev; /* ssim.c:52-53: (float)sqrt((double)var) */
  int32_t sum;  /* exact pixel sum (only meaningful on the integer path) */
} pos_stat;

/* ssim.c:3-27 + :52 for the w x h rectangle whose top-left pixel is (x, y) */
static pos_stat stats_at(const uint8_t *f, int stride, int x, int y, int w, int h, int lit) {
  pos_stat s;
  const float area = (float)(w * h);             /* ssim.c:12: int product -> float */
  int32_t isum = 0;
  if (lit || (int64_t)w * h > 65793) {
    float fs = 0;
    for (int oy = 0; oy < h; oy++)
      for (int ox = 0; ox < w; ox++) fs += f[(y + oy) * stride + x + ox];   /* ssim.c:9 */
    s.mean = fs / area;
    isum = (int32_t)fs;
  } else {
    for (int oy = 0; oy < h; oy++)
      for (int ox = 0; ox < w; ox++) isum += f[(y + oy) * stride + x + ox];
    s.mean = (float)isum / area;
  }
  s.sum = isum;
  float vs = 0;
  for (int oy = 0; oy < h; oy++)
    for (int ox = 0; ox < w; ox++) {
      const int p = f[(y + oy) * stride + x + ox];
      const float d = (float)p - s.mean;          /* ssim.c:22: int - float */
      vs += d * d;
    }
  const float var = vs / area;                    /* ssim.c:25 */
  s.stddev = (float)sqrt((double)var);            /* ssim.c:52 */
  return s;
}

/* ssim.c:29-41: means truncated to int by the call (ssim.c:54, ssim.h:12) */
static float cross_at(const uint8_t *ref, const uint8_t *cur, int stride, int cx, int cy, int bx, int by,
                      int w, int h, int imr, int imc, int lit) {
  const float area = (float)(w * h);
  if (lit || (int64_t)w * h > 258) {
    float fs = 0;
    for (int oy = 0; oy < h; oy++)
      for (int ox = 0; ox < w; ox++)
        fs += (ref[(cy + oy) * stride + cx + ox] - imr) * (cur[(by + oy) * stride + bx + ox] - imc);
    return fs / area;
  }
  int32_t is = 0;
  for (int oy = 0; oy < h; oy++)
    for (int ox = 0; ox < w; ox++)
      is += ((int)ref[(cy + oy) * stride + cx + ox] - imr) * ((int)cur[(by + oy) * stride + bx + ox] - imc);
  return (float)is / area;
}

/* ssim.c:44-60 from the two statistics and the cross term */
static float ssim_score(pos_stat r, pos_stat c, float cross) {
  const float C1 = 0.01, C2 = 0.09, C3 = 0.045;                           /* ssim.c:47 */
  const float lum = (2 * r.mean * c.mean + C1) / (r.mean * r.mean + c.mean * c.mean + C1);         /* :55 */
  const float con = (2 * r.stddev * c.stddev + C2) / (r.stddev * r.stddev + c.stddev * c.stddev + C2); /* :56 */
  const float str = (cross + C3) / (r.stddev * c.stddev + C3);           /* :57 */
  return lum * con * str;                                                 /* :58 */
}

typedef struct ssim_job {
  const uint8_t *cur, *ref;
  int W, H, B, R, begin, end, base, lit;
  const pos_stat *table; /* statistics of every full-size position of ref, or NULL */
  me_oracle_result *out;
} ssim_job;

static void ssim_block(const ssim_job *j, int i, me_oracle_result *out) {
  const int W = j->W, H = j->H, R = j->R;
  int x0, y0, w, h;
  me_oracle_block_geom(i, W, H, j->B, &x0, &y0, &w, &h);
  const int brx = x0 + w - 1, bry = y0 + h - 1;
  const int wx0 = (x0 - R) < 0 ? 0 : x0 - R;                 /* main_ssim.c:22 */
  const int wy0 = (y0 - R) < 0 ? 0 : y0 - R;                 /* :23 */
  const int wx1 = (brx + R) >= W ? W - 1 : brx + R;          /* :24 */
  const int wy1 = (bry + R) >= H ? H - 1 : bry + R;          /* :25 */
  const pos_stat cs = stats_at(j->cur, W, x0, y0, w, h, j->lit);   /* ssim.c:49,51,53 */
  const int imc = (int)cs.mean;
  const int full = j->table && w == j->B && h == j->B;
  float best = 0;                                             /* ssim.c:88 */
  int bx = 0, by = 0, found = 0;
  for (int y = wy0; y <= wy1 - h + 1; y++)                    /* ssim.c:98 */
    for (int x = wx0; x <= wx1 - w + 1; x++) {                /* ssim.c:99 */
      const pos_stat rs = full ? j->table[y * W + x] : stats_at(j->ref, W, x, y, w, h, j->lit);
      const float cross = cross_at(j->ref, j->cur, W, x, y, x0, y0, w, h, (int)rs.mean, imc, j->lit);
      const float s = ssim_score(rs, cs, cross);
      if (s > best) {                                         /* ssim.c:101 */
        best = s;
        bx = x - x0;                                          /* ssim.c:103 */
        by = y - y0;                                          /* ssim.c:104 */
        found = 1;
      }
    }
  out->mvx = bx;                                              /* main_ssim.c:27 */
  out->mvy = by;
  out->ssd = (uint32_t)found;
  out->score = best;
}

static void *ssim_job_main(void *p) {
  const ssim_job *j = (const ssim_job *)p;
  for (int i = j->begin; i < j->end; i++) ssim_block(j, i, &j->out[i - j->base]);
  return NULL;
}

typedef struct table_job {
  const uint8_t *ref;
  int W, H, B, y_begin, y_end, lit;
  pos_stat *table;
} table_job;

static void *table_job_main(void *p) {
  const table_job *t = (const table_job *)p;
  for (int y = t->y_begin; y < t->y_end; y++)
    for (int x = 0; x + t->B <= t->W; x++) t->table[y * t->W + x] = stats_at(t->ref, t->W, x, y, t->B, t->B, t->lit);
  return NULL;
}

int me_oracle_search_ssim(const uint8_t *cur, const uint8_t *ref, int width, int height, int blk_dim,
                          int extra_span, int blk_begin, int blk_end, int nthreads, me_oracle_result *out) {
  const int nb = me_oracle_num_blocks(width, height, blk_dim);
  if (!cur || !ref || !out || nb <= 0 || extra_span < 0) return -1;
  if (blk_begin < 0 || blk_end > nb || blk_begin > blk_end) return -1;
  const int n = blk_end - blk_begin;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  const int lit = ssim_literal();
  pthread_t th[256];

  /* tabulate the full-size statistics when the search is big enough to reuse them */
  pos_stat *table = NULL;
  const int64_t cands = (int64_t)n * (2 * extra_span + 1) * (2 * extra_span + 1);
  if (width >= blk_dim && height >= blk_dim && cands > 4 * (int64_t)width * height) {
    table = (pos_stat *)calloc((size_t)width * height, sizeof(pos_stat));
    if (!table) return -1;
    const int rows = height - blk_dim + 1;
    const int nt = nthreads > rows ? rows : nthreads;
    table_job tj[256];
    for (int t = 0; t < nt; t++) {
      table_job v = {ref, width, height, blk_dim, (int)((int64_t)rows * t / nt), (int)((int64_t)rows * (t + 1) / nt),
                     lit, table};
      tj[t] = v;
      pthread_create(&th[t], NULL, table_job_main, &tj[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  }

  const int nt = nthreads > n ? (n > 0 ? n : 1) : nthreads;
  ssim_job jobs[256];
  for (int t = 0; t < nt; t++) {
    ssim_job j = {cur, ref, width, height, blk_dim, extra_span,
                  blk_begin + (int)((int64_t)n * t / nt), blk_begin + (int)((int64_t)n * (t + 1) / nt),
                  blk_begin, lit, table, out};
    jobs[t] = j;
    if (nt == 1) ssim_job_main(&jobs[0]);
    else pthread_create(&th[t], NULL, ssim_job_main, &jobs[t]);
  }
  if (nt > 1)
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  free(table);
  return 0;
}

/* The two numbers main_ssim.c:88-95 prints: float-accumulated squared errors of the whole
 * frame, divided by the pixel count (int -> float). */
void me_oracle_ssim_frame_scores(const uint8_t *cur, const uint8_t *ref, const uint8_t *mc, int n,
                                 float *original_score, float *compensated_score) {
  float motionCompScore = 0.0, originalScore = 0.0;
  for (int i = 0; i < n; i++) {
    const int a = (int)mc[i] - (int)cur[i], b = (int)cur[i] - (int)ref[i];
    motionCompScore += a * a;                               /* main_ssim.c:91 */
    originalScore += b * b;                                 /* main_ssim.c:92 */
  }
  *original_score = originalScore / n;                      /* main_ssim.c:94 */
  *compensated_score = motionCompScore / n;
}
