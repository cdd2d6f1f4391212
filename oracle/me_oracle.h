/*
 * me_oracle.h -- CPU restatement of the reference full-search MSE path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (1) against the reference's own golden outputs
 *       (results/cpu/foreman/output_4_15.yuv, output_4_7.yuv md5s and the
 *        logged PSNR 31.816000 / 31.750712), and
 *   (2) block by block (mv + score bits) against the UNMODIFIED reference
 *       functions compiled from /root/reference into oracle/_ref/libme_ref.so
 *       (oracle/ref_harness.c).
 *
 * Every function cites the reference file:line it restates
 * (paths relative to the reference checkout).
 */
#ifndef ME_ORACLE_H
#define ME_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-block result.  mvx/mvy/score are what the reference computes
 * (src/cpu/main.c:56-60,79); ssd is the exact integer sum of squared
 * differences of the winning candidate (the reference only holds it as the
 * float numerator `sum`, main.c:19-27). */
typedef struct me_oracle_result {
  int32_t  mvx;
  int32_t  mvy;
  uint32_t ssd;
  float    score;
} me_oracle_result;

/* Number of blocks = ceil(W/B)*ceil(H/B)  (src/common/prediction_frame.c:9-12). */
int me_oracle_num_blocks(int width, int height, int blk_dim);

/* Geometry of block i (src/common/prediction_frame.c:14-23, block.c:3-13). */
void me_oracle_block_geom(int i, int width, int height, int blk_dim,
                          int *x0, int *y0, int *w, int *h);

/* Full search for blocks [blk_begin, blk_end) of one frame pair.
 * cur/ref are 8-bit luma planes, row-major, stride == width.
 * out[i - blk_begin] receives block i.  Restates main.c:18-82.
 * nthreads > 1 splits the block range over pthreads (blocks are independent,
 * main.c:152-154).  Returns 0, or -1 on bad arguments. */
int me_oracle_search(const uint8_t *cur, const uint8_t *ref,
                     int width, int height, int blk_dim, int extra_span,
                     int blk_begin, int blk_end, int nthreads,
                     me_oracle_result *out);

/* Exact pixel-compare count of one frame: number of
 * (cur-ref)^2 terms the reference evaluates. */
uint64_t me_oracle_pixel_compares(int width, int height, int blk_dim, int extra_span);
/* Number of candidates over the frame. */
uint64_t me_oracle_candidates(int width, int height, int blk_dim, int extra_span);

/* Post-search stage on 8-bit planes (src/common/utils.c:94-134): writes the
 * motion-compensated plane; returns -1 if an MV points outside the frame
 * (the reference silently skips such pixels, utils.c:122; full search never
 * produces them). */
int me_oracle_motion_compensate(const uint8_t *ref, int width, int height, int blk_dim,
                                const me_oracle_result *res, uint8_t *mc);
/* |a-b| per pixel (utils.c:94-100). */
void me_oracle_frame_diff(const uint8_t *a, const uint8_t *b, int n, uint8_t *out);
/* PSNR with peak = max pixel of both frames, 99.0 when identical (utils.c:137-164). */
double me_oracle_psnr(const uint8_t *a, const uint8_t *b, int width, int height);

/* ---- SSIM-cost full search (oracle/me_oracle_ssim.c; PINNED against the unmodified
 * src/cpu/main_ssim.c + src/common/ssim.c).  Same arguments as me_oracle_search.
 * out[].score = best SSIM (0 when no candidate scored above 0), out[].ssd = 1 when some
 * candidate scored above 0, else 0 (the reference then leaves the MV uninitialised; (0,0) here). */
int me_oracle_search_ssim(const uint8_t *cur, const uint8_t *ref,
                          int width, int height, int blk_dim, int extra_span,
                          int blk_begin, int blk_end, int nthreads,
                          me_oracle_result *out);
/* "Original Score" / "Compensated Score" of main_ssim.c:88-95 (float accumulation). */
void me_oracle_ssim_frame_scores(const uint8_t *cur, const uint8_t *ref, const uint8_t *mc, int n,
                                 float *original_score, float *compensated_score);

/* ---- fast searches (oracle/me_oracle_fast.c; PARITY UNPINNED: absent from the reference,
 * defined by that file).  algo 1 = three-step, 2 = diamond; MSE cost of main.c:18-27. */
int me_oracle_search_fast(const uint8_t *cur, const uint8_t *ref, int width, int height, int blk_dim,
                          int extra_span, int algo, int blk_begin, int blk_end, int nthreads,
                          me_oracle_result *out, uint64_t *evaluated);
int me_oracle_tss_first_step(int extra_span);

#ifdef __cplusplus
}
#endif
#endif
