/*
 * mes_b200_ssim -- drop-in command line for the reference SSIM program (plain C).
 * Same argv, same stdout lines and the same output file as src/cpu/main_ssim.c:
 *   argv    <cur> <ref> <outdir> [blk=16] [span=7] [W=3840] [H=2160]      (main_ssim.c:33-44)
 *   stdout  parameter banner, "Original Score: %.4f, Compensated Score: %.4f",
 *           "Output file dimensions: (W x 5H)"                          (main_ssim.c:45-46,95,98)
 *   file    <outdir>/output_<blk>_<span>.yuv = 5 stacked 8-bit planes     (main_ssim.c:53,80-99)
 * The search loop (main_ssim.c:67-77) is ONE call into the CUDA library,
 * me_b200_search_ssim_scores().  Additionally <outdir>/mv_ssim_<blk>_<span>.txt holds the
 * per-block field: idx x0 y0 w h mvx mvy found score-bits(hex).
 * No CPU search exists in this program: without a GPU it reports the error and exits 2.
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
  const int blkDim = argc > 4 ? atoi(argv[4]) : 16;
  const int extraSpan = argc > 5 ? atoi(argv[5]) : 7;
  const int W = argc > 6 ? atoi(argv[6]) : 3840;
  const int H = argc > 7 ? atoi(argv[7]) : 2160;
  printf("[\n  Current Frame: %s\n  Reference Frame: %s\n  Output Dir: %s\n  BlkDim: %d\n  ExtraSpan: %d\n  FrameWidth: %d\n  FrameHeight: %d\n]\n",
         curName, refName, outDir, blkDim, extraSpan, W, H);
  if (blkDim <= 0 || extraSpan < 0 || W <= 0 || H <= 0) {
    printf("Error: invalid parameters\n");
    return 2;
  }

  const int numElems = W * H;
  int *cur = (int *)malloc(sizeof(int) * (size_t)numElems);
  int *ref = (int *)malloc(sizeof(int) * (size_t)numElems);
  if (!cur || !ref) return 2;
  if (!yuvReadFrame(curName, cur, numElems)) exit(1);
  if (!yuvReadFrame(refName, ref, numElems)) exit(1);

  predictionFrame p;
  createPredictionFrame(&p, cur, W, H, blkDim);
  float *scores = (float *)malloc(sizeof(float) * (size_t)p.num_blks);
  uint32_t *found = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)p.num_blks);
  int rc = me_b200_search_ssim_scores(&p, ref, extraSpan, scores, found);
  if (rc != ME_OK) {
    fprintf(stderr, "mes_b200_ssim: search failed: %s (%s)\n", me_b200_strerror(rc), me_b200_last_error(NULL));
    return 2;
  }

  int *out = (int *)calloc((size_t)numElems * 5, sizeof(int));
  memcpy(out, ref, sizeof(int) * (size_t)numElems);
  memcpy(out + numElems, cur, sizeof(int) * (size_t)numElems);
  if (!motionCompensatedFrame(out + 2 * numElems, p, ref)) {
    printf("Error: Trying to create compensation frame without best match\n");
    exit(0);
  }
  frameDiff(out + 3 * numElems, ref, cur, numElems);
  frameDiff(out + 4 * numElems, out + 2 * numElems, cur, numElems);

  /* main_ssim.c:88-95: squared errors accumulated in float, in pixel order */
  float motionCompScore = 0.0, originalScore = 0.0;
  for (int i = 0; i < numElems; i++) {
    motionCompScore += (out[numElems * 2 + i] - cur[i]) * (out[numElems * 2 + i] - cur[i]);
    originalScore += (cur[i] - ref[i]) * (cur[i] - ref[i]);
  }
  printf("Original Score: %.4f, Compensated Score: %.4f\n", originalScore / numElems, motionCompScore / numElems);
  printf("Output file dimensions: (%d x %d)\n", W, 5 * H);
  char name[4096];
  snprintf(name, sizeof name, "%s/output_%d_%d.yuv", outDir, blkDim, extraSpan);
  yuvWriteFrame(name, out, numElems * 5);

  snprintf(name, sizeof name, "%s/mv_ssim_%d_%d.txt", outDir, blkDim, extraSpan);
  FILE *f = fopen(name, "w");
  if (f) {
    for (int i = 0; i < p.num_blks; i++) {
      const block *b = &p.blks[i];
      uint32_t bits;
      memcpy(&bits, &scores[i], 4);
      fprintf(f, "%d %d %d %d %d %d %d %u %08x\n", i, b->top_left_x, b->top_left_y, b->width, b->height,
              b->motion_vectorX, b->motion_vectorY, found[i], bits);
    }
    fclose(f);
  }
  free(out); free(scores); free(found); free(p.blks); free(cur); free(ref);
  return 0;
}
