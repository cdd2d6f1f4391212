/*
 * me_common.h -- host data model of the drop-in, binary-identical to the
 * reference's src/common interface so existing callers keep compiling:
 *
 *   block            <- src/common/block.h:6-19            (11 x int = 44 B, AoS)
 *   predictionFrame  <- src/common/prediction_frame.h:8-16 (frame ptr, dims, grid)
 *   createBlk                <- src/common/block.c:3-13
 *   createPredictionFrame    <- src/common/prediction_frame.c:3-25
 *   yuvReadFrame/yuvWriteFrame      <- src/common/utils.c:61-92
 *   frameDiff/motionCompensatedFrame/imagePSNR/getTimeStamp
 *                                   <- src/common/utils.c:23-27, 94-164
 *
 * Same names, argument meaning and return conventions (1 = ok, 0 = failure for
 * the I/O functions).  Deliberate differences, all on error paths only:
 *   - yuvReadFrame returns 0 when the file cannot be opened (the reference
 *     dereferences a NULL FILE*, utils.c:62-64) and closes the file.
 *   - motionCompensatedFrame returns 0 instead of exit(0) when a block has no
 *     motion vector (utils.c:105-108); the drop-in CLI turns that into the
 *     reference's message + exit.
 * Implemented in plain C in motionestimation_b200/host/.
 *
 * Inside the reference tree: when the reference's own
 * ../common/block.h, prediction_frame.h and utils.h have been included first (their
 * include guards BLOCK_H / PREDICTION_FRAME_H / UTILS_H are visible), or when
 * ME_B200_REFERENCE_TYPES is defined, this header declares nothing twice: the two
 * structs and the src/common functions are the reference's own.
 */
#ifndef ME_COMMON_H
#define ME_COMMON_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(BLOCK_H) && defined(PREDICTION_FRAME_H) && !defined(ME_B200_REFERENCE_TYPES)
#define ME_B200_REFERENCE_TYPES 1
#endif

#ifndef ME_B200_REFERENCE_TYPES
typedef struct block {
  int idx_x;
  int idx_y;
  int top_left_x;
  int top_left_y;
  int bottom_right_x; /* inclusive */
  int bottom_right_y; /* inclusive */
  int width;
  int height;
  int is_best_match_found;
  int motion_vectorX;
  int motion_vectorY;
} block;

typedef struct predictionFrame {
  int *frame; /* current frame, one int per pixel, values 0..255 */
  int width;
  int height;
  int blk_dim;
  int num_blks;
  block *blks; /* raster order: i = by * ceil(W/B) + bx */
} predictionFrame;

void createBlk(block *blk, int idxX, int idxY, int topLeftX, int topLeftY, int width, int height);
void createPredictionFrame(predictionFrame *pf, int *frame, int width, int height, int blkDim);

double getTimeStamp(void);
int yuvReadFrame(const char *file_name, int *const target_buffer, int numElems);
int yuvWriteFrame(const char *file_name, const int *const data_buffer, int numElems);
void frameDiff(int *diffFrame, const int *frameA, const int *frameB, int numElems);
int motionCompensatedFrame(int *motionCompFrame, predictionFrame pf, const int *ref_frame);
double imagePSNR(const int *frame1, const int *frame2, int x, int y);
#endif /* ME_B200_REFERENCE_TYPES */

/* u8 ingest without the int detour: reads numElems
 * bytes of the first luma plane straight into a byte buffer. 1 ok / 0 fail. */
int yuvReadFrameU8(const char *file_name, unsigned char *target_buffer, int numElems);

#ifdef __cplusplus
}
#endif
#endif
