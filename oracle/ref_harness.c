/*
 * ref_harness.c -- thin harness around the UNMODIFIED reference CPU path.
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into
 * oracle/_ref/libme_ref.so, compiling the reference sources where they lie
 * under $(REF) (= /root/reference); no reference source is copied into this
 * repository.  The reference's own main() is renamed at compile time
 * (-Dmain=ref_main) so its non-static functions
 *   findBestBlkMse      src/cpu/main.c:67
 *   runFindBestBlkMse   src/cpu/main.c:101
 *   createRunConfig     src/cpu/main.c:93
 * and the src/common functions can be called per block.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define main ref_main
#include REF_MAIN_C /* "<ref>/src/cpu/main.c", given on the command line */
#undef main

typedef struct ref_result {
  int32_t  mvx;
  int32_t  mvy;
  uint32_t ssd;   /* not produced by the reference; left 0 */
  float    score; /* return value of findBestBlkMse, main.c:81 */
} ref_result;

static int *widen(const uint8_t *src, int n) {
  int *dst = (int *)malloc(sizeof(int) * (size_t)n);
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
    float s = findBestBlkMse(*j->p, j->ref, &j->p->blks[i], j->R);
    ref_result *o = &j->out[i - j->base];
    o->mvx = j->p->blks[i].motion_vectorX;
    o->mvy = j->p->blks[i].motion_vectorY;
    o->ssd = 0;
    o->score = s;
  }
  return NULL;
}

/* Per-block float scores + MVs for blocks [begin,end), calling the reference's
 * findBestBlkMse directly from nthreads plain pthreads. */
int ref_search_blocks(const uint8_t *cur, const uint8_t *ref, int W, int H, int B, int R,
                      int begin, int end, int nthreads, ref_result *out) {
  int n = W * H;
  int *c = widen(cur, n), *r = widen(ref, n);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  if (begin < 0 || end > p.num_blks || begin > end) return -1;
  int cnt = end - begin;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > cnt) nthreads = cnt > 0 ? cnt : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  span_job *jobs = (span_job *)malloc(sizeof(span_job) * (size_t)nthreads);
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

/* The reference's own timed region, main.c:144-158, on blocks [begin,end):
 * thpool_init(pool_threads) (the reference uses 100), one job per block via
 * runFindBestBlkMse, thpool_wait; returns the seconds between the two
 * getTimeStamp() calls exactly as main.c:151,157 measure them.  MVs are
 * returned through out (score = the int-truncated config->score, main.c:104-105). */
double ref_search_pool(const uint8_t *cur, const uint8_t *ref, int W, int H, int B, int R,
                       int begin, int end, int pool_threads, ref_result *out) {
  int n = W * H;
  int *c = widen(cur, n), *r = widen(ref, n);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  if (begin < 0 || end > p.num_blks || begin > end) return -1.0;
  int cnt = end - begin;
  threadpool thpool = thpool_init(pool_threads);
  runConfig **configs = (runConfig **)malloc(sizeof(runConfig *) * (size_t)(cnt > 0 ? cnt : 1));
  for (int i = 0; i < cnt; i++) {
    runConfig *config = (runConfig *)malloc(sizeof(runConfig));
    createRunConfig(config, &p, r, &p.blks[begin + i], R);
    configs[i] = config;
  }
  double t0 = getTimeStamp();
  for (int i = 0; i < cnt; i++) thpool_add_work(thpool, runFindBestBlkMse, configs[i]);
  thpool_wait(thpool);
  double t1 = getTimeStamp();
  thpool_destroy(thpool);
  for (int i = 0; i < cnt; i++) {
    if (out) {
      out[i].mvx = p.blks[begin + i].motion_vectorX;
      out[i].mvy = p.blks[begin + i].motion_vectorY;
      out[i].ssd = 0;
      out[i].score = (float)configs[i]->score;
    }
    free(configs[i]);
  }
  free(configs); free(p.blks); free(c); free(r);
  return t1 - t0;
}

/* Reference post-processing, main.c:160-171: builds the 5 stacked planes
 * (ref, cur, motion-compensated, |ref-cur|, |mc-cur|) from the given MV field
 * with the reference's own motionCompensatedFrame/frameDiff/imagePSNR and
 * narrows them like yuvWriteFrame (utils.c:55-59).  Returns the PSNR. */
double ref_postprocess(const uint8_t *cur, const uint8_t *ref, int W, int H, int B,
                       const int32_t *mvx, const int32_t *mvy, uint8_t *out5) {
  int n = W * H;
  int *c = widen(cur, n), *r = widen(ref, n);
  predictionFrame p;
  createPredictionFrame(&p, c, W, H, B);
  for (int i = 0; i < p.num_blks; i++)
    populateBlkMotionVector(&p.blks[i], mvx[i], mvy[i]);
  int *o = (int *)calloc((size_t)n * 5, sizeof(int));
  memcpy(o, r, sizeof(int) * (size_t)n);
  memcpy(&o[n], c, sizeof(int) * (size_t)n);
  motionCompensatedFrame(&o[n * 2], p, r);
  frameDiff(&o[n * 3], r, c, n);
  frameDiff(&o[n * 4], &o[n * 2], c, n);
  double psnr = imagePSNR(&o[n * 2], c, W, H);
  for (int i = 0; i < n * 5; i++) out5[i] = (uint8_t)o[i];
  free(o); free(p.blks); free(c); free(r);
  return psnr;
}

int ref_sizeof_block(void) { return (int)sizeof(block); }
int ref_sizeof_prediction_frame(void) { return (int)sizeof(predictionFrame); }
