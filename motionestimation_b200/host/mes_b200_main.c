/*
 * mes_b200 -- drop-in command line for the reference CPU program (plain C).
 * Same argv, same stdout lines and the same output file as src/cpu/main.c:
 *   argv    <cur> <ref> <outdir> [blk=8] [span=12] [W=352] [H=288]   (main.c:110-120)
 *   stdout  parameter banner, "PSNR: %.6f", "Output file dimensions",
 *           "Computation time: %.lf ms", "PSNR: %.lf "               (main.c:121-122,171-178)
 *   file    <outdir>/output_<blk>_<span>.yuv = 5 stacked 8-bit planes (main.c:129,161-175)
 * The search itself (main.c:144-158) is ONE call into the CUDA library,
 * me_b200_search_scores().  In addition the per-block field the reference keeps
 * only in memory is written to <outdir>/mv_<blk>_<span>.txt:
 *   idx x0 y0 w h mvx mvy ssd score-bits(hex)
 * Extra option (environment only, so the argv stays the reference's): ME_B200_SEARCH=tss or
 * =diamond runs the three-step / diamond pattern instead of the exhaustive scan (not in the
 * reference; see include/me_b200.h).
 * There is no CPU search in this program: without a GPU it reports the error
 * and exits 2.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "me_b200.h"

int main(int argc, char *argv[]) {
  if (argc < 4) {
    printf("Error: wrong number of argument. Usage: <current_frame> <reference_frame> <output_dir> [<blk_dim>] [<extra_span>] [<width>] [<height>]\n");
    exit(0);
  }
  const char *curName = argv[1];
  const char *refName = argv[2];
  const char *outDir = argv[3];
  const int blkDim = argc > 4 ? atoi(argv[4]) : 8;
  const int extraSpan = argc > 5 ? atoi(argv[5]) : 12;
  const int W = argc > 6 ? atoi(argv[6]) : 352;
  const int H = argc > 7 ? atoi(argv[7]) : 288;
  printf("[\n  Current Frame: %s\n  Reference Frame: %s\n  Output Dir: %s\n  BlkDim: %d\n  ExtraSpan: %d\n  FrameWidth: %d\n  FrameHeight: %d\n]\n",
         curName, refName, outDir, blkDim, extraSpan, W, H);
  if (blkDim <= 0 || extraSpan < 0 || W <= 0 || H <= 0) {
    printf("Error: invalid parameters\n");
    return 2;
  }

  const int n = W * H;
  int *cur = (int *)malloc(sizeof(int) * (size_t)n);
  int *ref = (int *)malloc(sizeof(int) * (size_t)n);
  if (!cur || !ref) return 2;
  if (!yuvReadFrame(curName, cur, n)) exit(1);
  if (!yuvReadFrame(refName, ref, n)) exit(1);

  predictionFrame p;
  createPredictionFrame(&p, cur, W, H, blkDim);
  float *scores = (float *)malloc(sizeof(float) * (size_t)p.num_blks);
  uint32_t *ssd = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)p.num_blks);

  int pattern = ME_SEARCH_FULL;
  const char *se = getenv("ME_B200_SEARCH");
  if (se && !strcmp(se, "tss")) pattern = ME_SEARCH_THREE_STEP;
  if (se && !strcmp(se, "diamond")) pattern = ME_SEARCH_DIAMOND;

  /* warm the context (device init, buffers) outside the timed region, as the
   * reference keeps thpool_init outside its timestamps (main.c:144,151) */
  int rc = pattern == ME_SEARCH_FULL ? me_b200_search_scores(&p, ref, extraSpan, scores, ssd)
                                     : me_b200_search_fast(&p, ref, extraSpan, pattern, scores, ssd);
  if (rc == ME_OK) {
    double t0 = getTimeStamp();
    rc = pattern == ME_SEARCH_FULL ? me_b200_search_scores(&p, ref, extraSpan, scores, ssd)
                                   : me_b200_search_fast(&p, ref, extraSpan, pattern, scores, ssd);
    double t1 = getTimeStamp();
    if (rc == ME_OK) {
      int *out = (int *)calloc((size_t)n * 5, sizeof(int));
      memcpy(out, ref, sizeof(int) * (size_t)n);
      memcpy(out + n, cur, sizeof(int) * (size_t)n);
      if (!motionCompensatedFrame(out + 2 * n, p, ref)) {
        printf("Error: Trying to create compensation frame without best match\n");
        exit(0);
      }
      frameDiff(out + 3 * n, ref, cur, n);
      frameDiff(out + 4 * n, out + 2 * n, cur, n);
      printf("PSNR: %.6f\n", imagePSNR(out + 2 * n, cur, W, H));
      printf("Output file dimensions: (%d x %d)\n", W, 5 * H);
      char name[4096];
      snprintf(name, sizeof name, "%s/output_%d_%d.yuv", outDir, blkDim, extraSpan);
      yuvWriteFrame(name, out, n * 5);
      printf("Computation time: %.lf ms\n", (t1 - t0) * 1000);
      printf("PSNR: %.lf \n", imagePSNR(out + 2 * n, cur, W, H));

      snprintf(name, sizeof name, "%s/mv_%d_%d.txt", outDir, blkDim, extraSpan);
      FILE *f = fopen(name, "w");
      if (f) {
        for (int i = 0; i < p.num_blks; i++) {
          const block *b = &p.blks[i];
          uint32_t bits;
          memcpy(&bits, &scores[i], 4);
          fprintf(f, "%d %d %d %d %d %d %d %u %08x\n", i, b->top_left_x, b->top_left_y, b->width,
                  b->height, b->motion_vectorX, b->motion_vectorY, ssd[i], bits);
        }
        fclose(f);
      }
      free(out);
    }
  }
  if (rc != ME_OK) {
    fprintf(stderr, "mes_b200: search failed: %s (%s)\n", me_b200_strerror(rc), me_b200_last_error(NULL));
    return 2;
  }
  free(scores); free(ssd); free(p.blks); free(cur); free(ref);
  return 0;
}
