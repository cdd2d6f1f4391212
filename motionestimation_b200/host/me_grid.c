/*
 * me_grid.c -- block grid of the drop-in host layer (plain C).
 * Behaviour of reference src/common/block.c:3-13 and
 * src/common/prediction_frame.c:3-25: raster-order tiling, ceil(W/B) x
 * ceil(H/B) blocks, partial blocks on the right/bottom edges keep their
 * reduced extent, corners stored inclusive.
 */
#include <stdlib.h>
#include "me_common.h"

void createBlk(block *blk, int idxX, int idxY, int topLeftX, int topLeftY, int width, int height) {
  blk->idx_x = idxX;
  blk->idx_y = idxY;
  blk->top_left_x = topLeftX;
  blk->top_left_y = topLeftY;
  blk->bottom_right_x = topLeftX + width - 1;
  blk->bottom_right_y = topLeftY + height - 1;
  blk->width = width;
  blk->height = height;
  /* The reference leaves these two uninitialised (block.c:3-13 sets only
   * motion_vectorY = -1000); a defined "no match yet" state is a superset. */
  blk->is_best_match_found = 0;
  blk->motion_vectorX = 0;
  blk->motion_vectorY = -1000;
}

void createPredictionFrame(predictionFrame *pf, int *frame, int width, int height, int blkDim) {
  const int nbx = (width + blkDim - 1) / blkDim;
  const int nby = (height + blkDim - 1) / blkDim;
  pf->frame = frame;
  pf->width = width;
  pf->height = height;
  pf->blk_dim = blkDim;
  pf->num_blks = nbx * nby;
  pf->blks = (block *)malloc(sizeof(block) * (size_t)pf->num_blks);
  if (!pf->blks) {
    pf->num_blks = 0;
    return;
  }
  int i = 0;
  for (int by = 0; by < nby; by++) {
    const int y0 = by * blkDim;
    const int h = (y0 + blkDim < height) ? blkDim : height - y0;
    for (int bx = 0; bx < nbx; bx++, i++) {
      const int x0 = bx * blkDim;
      const int w = (x0 + blkDim < width) ? blkDim : width - x0;
      createBlk(&pf->blks[i], bx, by, x0, y0, w, h);
    }
  }
}
