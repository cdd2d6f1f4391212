/*
 * me_post.c -- host post-search stage of the drop-in (plain C).
 * Behaviour of reference src/common/utils.c:94-100 (frameDiff), :102-134
 * (motionCompensatedFrame) and :137-164 (imagePSNR: peak = largest pixel of
 * either frame, not 255; 99.0 for identical frames; double arithmetic).
 */
#include <math.h>
#include <stdlib.h>
#include "me_common.h"

void frameDiff(int *diffFrame, const int *frameA, const int *frameB, int numElems) {
  for (int i = 0; i < numElems; i++) {
    const int d = frameA[i] - frameB[i];
    diffFrame[i] = d < 0 ? -d : d;
  }
}

/* Copies, for every block, the reference pixels the motion vector points at.
 * Pixels whose source falls outside the frame are left untouched, as in the
 * reference (utils.c:122).  Returns 1, or 0 if some block carries no vector
 * (the reference prints and exit(0)s there, utils.c:105-108). */
int motionCompensatedFrame(int *motionCompFrame, predictionFrame pf, const int *ref_frame) {
  for (int i = 0; i < pf.num_blks; i++) {
    const block *b = &pf.blks[i];
    if (b->is_best_match_found != 1) return 0;
    const int sx = b->top_left_x + b->motion_vectorX;
    const int sy = b->top_left_y + b->motion_vectorY;
    for (int oy = 0; oy < b->height; oy++) {
      const int y = sy + oy;
      if (y < 0 || y >= pf.height) continue;
      for (int ox = 0; ox < b->width; ox++) {
        const int x = sx + ox;
        if (x < 0 || x >= pf.width) continue;
        motionCompFrame[(b->top_left_y + oy) * pf.width + b->top_left_x + ox] =
            ref_frame[y * pf.width + x];
      }
    }
  }
  return 1;
}

double imagePSNR(const int *frame1, const int *frame2, int x, int y) {
  const int n = x * y;
  double sq = 0.0;
  int peak = 0;
  for (int i = 0; i < n; i++) {
    if (frame1[i] > peak) peak = frame1[i];
    if (frame2[i] > peak) peak = frame2[i];
    const double d = abs(frame1[i] - frame2[i]);
    sq += d * d;
  }
  const double mse = sq / n;
  if (mse == 0) return 99.0;
  return 20 * log10(peak) - 10 * log10(mse);
}
