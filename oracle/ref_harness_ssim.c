/*
 * ref_harness_ssim.c -- thin harness around the UNMODIFIED reference SSIM search.
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into
 * oracle/_ref/libme_ref_ssim.so from the reference sources where they lie under
 * $(REF): src/cpu/main_ssim.c (its main() renamed at compile time), src/common/ssim.c
 * and src/common/{block,prediction_frame,utils}.c.  Nothing is copied.
 *
 * Called per block:
 *   findBestBlkSSIM    src/cpu/main_ssim.c:16   (window clamp, MV write-back)
 *   findBestMatchSSIM  src/common/ssim.c:83     (y-major/x-minor scan, strict '>' against 0)
 *   computeSSIM        src/common/ssim.c:44
 *
 * One known defect of the reference is made deterministic here: when no candidate
 * scores above 0 the scan never writes result[1], result[2] (ssim.c:88-103), so
 * findBestBlkSSIM (main_ssim.c:26-27) turns uninitialised heap bytes into the motion
 * vector.  The library is linked with -Wl,--wrap=malloc and the wrapper below returns
 * zeroed memory, which is also what a fresh heap gives the stand-alone program in
 * practice: such blocks report MV (0, 0) and score 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

void *__wrap_malloc(size_t n) { return calloc(1, n ? n : 1); }

#define main ref_ssim_main
#include REF_MAIN_SSIM_C /* "<ref>/src/cpu/main_ssim.c", given on the command line */
#undef main

typedef struct ref_result {
  int32_t  mvx;
  int32_t  mvy;
  uint32_t ssd;   /* 1 when some candidate scored above 0, else 0 */
  float    score; /* return value of findBestBlkSSIM, main_ssim.c:29 */
} ref_result;

static int *widen(const uint8_t *src, int n) {
  int *dst = (int *)calloc((size_t)n, sizeof(int));
  for (int i = 0; i < n; i++) dst[i] = (int)src[i]; /* as utils.c:49-53 */
  return dst;
}

typedef struct span_job {
  predictionFrame *p;
  int *ref;
  int R, begin, end, base;
  ref_result *out;
} span_job;

static void *span_main(void *a) {
  span_job *j = (span_job *)a;
  for (int i = j->begin; i < j->end; i++) {
    float s = findBestBlkSSIM(*j->p, j->ref, &j->p->blks[i], j->R);
    ref_result *o = &j->out[i - j->base];
    o->mvx = j->p->blks[i].motion_vectorX;
    o->mvy = j->p->blks[i].motion_vectorY;
    o->ssd = s > 0 ? 1u : 0u;
    o->score = s;
  }
  return NULL;
}

/* Blocks [begin,end) through the reference's findBestBlkSSIM.  The reference runs them
 * sequentially (main_ssim.c:67-77); blocks are independent, so nthreads > 1 only splits
 * the range over plain pthreads to keep the tests short. */
int ref_ssim_search_blocks(const uint8_t *cur, const uint8_t *ref, int W, int H, int B, int R,
                           int begin, int end, int nthreads, ref_result *out) {
  int n = W * H;
  int *c = widen(cur, n), *r = widen(ref, n);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  if (begin < 0 || end > p.num_blks || begin > end) return -1;
  int cnt = end - begin;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > cnt) nthreads = cnt > 0 ? cnt : 1;
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  span_job *jobs = (span_job *)calloc((size_t)nthreads, sizeof(span_job));
  for (int t = 0; t < nthreads; t++) {
    span_job j = {&p, r, R, begin + (int)((int64_t)cnt * t / nthreads),
                  begin + (int)((int64_t)cnt * (t + 1) / nthreads), begin, out};
    jobs[t] = j;
    pthread_create(&th[t], NULL, span_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs); free(p.blks); free(c); free(r);
  return 0;
}

/* One candidate's score, straight from computeSSIM (ssim.c:44-60). */
float ref_ssim_score(const uint8_t *cur, const uint8_t *ref, int W, int H, int B, int blk_index,
                     int cand_x, int cand_y) {
  int n = W * H;
  int *c = widen(cur, n), *r = widen(ref, n);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  float s = computeSSIM(r, cand_x, cand_y, c, p.blks[blk_index], W);
  free(p.blks); free(c); free(r);
  return s;
}

/* The tail of main_ssim.c:80-95 on a given MV field: the 5 stacked planes and the two
 * float-accumulated scores it prints ("Original Score", "Compensated Score").  The
 * accumulation loop is restated from main_ssim.c:88-95 (it lives inside main()). */
void ref_ssim_postprocess(const uint8_t *cur, const uint8_t *ref, int W, int H, int B,
                          const int32_t *mvx, const int32_t *mvy, uint8_t *out5,
                          float *original_score, float *compensated_score) {
  int numElems = W * H;
  int *c = widen(cur, numElems), *r = widen(ref, numElems);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  for (int i = 0; i < p.num_blks; i++) populateBlkMotionVector(&p.blks[i], mvx[i], mvy[i]);
  int *o = (int *)calloc((size_t)numElems * 5, sizeof(int));
  memcpy(o, r, sizeof(int) * (size_t)numElems);
  memcpy(&o[numElems], c, sizeof(int) * (size_t)numElems);
  motionCompensatedFrame(&o[numElems * 2], p, r);
  frameDiff(&o[numElems * 3], r, c, numElems);
  frameDiff(&o[numElems * 4], &o[numElems * 2], c, numElems);
  float motionCompScore = 0.0, originalScore = 0.0;
  for (int i = 0; i < numElems; i++) {
    motionCompScore += (o[numElems * 2 + i] - c[i]) * (o[numElems * 2 + i] - c[i]);
    originalScore += (c[i] - r[i]) * (c[i] - r[i]);
  }
  *original_score = originalScore / numElems;
  *compensated_score = motionCompScore / numElems;
  for (int i = 0; i < numElems * 5; i++) out5[i] = (uint8_t)o[i];
  free(o); free(p.blks); free(c); free(r);
}
