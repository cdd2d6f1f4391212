/*
 * me_b200.h -- C ABI of the B200-native full-search block-matching estimator.
 *
 * This is the drop-in boundary for the reference's CPU search path.  The
 * reference has no plugin/FFI layer: its search is the timed region of
 * src/cpu/main.c:144-158 (thread pool -> runFindBestBlkMse -> findBestBlkMse ->
 * findBestMatchMse -> computeMse, main.c:18-107).  Each entry point below names
 * the reference interface it replaces.  Everything is extern "C", plain
 * pointers and sizes, callable from C99; the implementation is hand-written
 * sm_100a CUDA (motionestimation_b200/csrc/).  There is NO CPU fallback: every
 * compute entry point returns ME_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Semantics (bit-exact with the reference CPU path):
 *   blocks   nbx=ceil(W/B), nby=ceil(H/B), i=by*nbx+bx, partial edge blocks kept
 *            (src/common/prediction_frame.c:9-23)
 *   window   [x0-R, x0+w-1+R] x [y0-R, y0+h-1+R] clamped to the frame (main.c:69-76)
 *   cost     SSD=sum (cur-ref)^2 accumulated in float, score=sum/(w*h)   (main.c:18-27)
 *   winner   first strict minimum of score in y-major, x-minor order     (main.c:53-62)
 *   output   mvx=x-x0, mvy=y-y0 (main.c:58-59,79), score (main.c:81), plus the exact
 *            integer SSD of the winner (the reference keeps it only as the float sum)
 */
#ifndef ME_B200_H
#define ME_B200_H

#include <stddef.h>
#include <stdint.h>
#include "me_common.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ME_B200_ABI_VERSION 3 /* 2: + cost / search modes, SSIM and fast drop-ins, peer fields;
                               * 3: + me_b200_last_kernel / _fallback_launches, batched post stage,
                               *    me_b200_host_alloc_ex, me_b200_set_ingest_helper (all additive) */

/* return codes: 0 ok, negative error.  The library never prints or exits
 * (the reference printf+exit()s, main.c:110-113,134-139; utils.c:105-108). */
#define ME_OK               0
#define ME_ERR_INVALID_ARG -1 /* NULL pointer, non-positive dimension, bad slot ...            */
#define ME_ERR_UNSUPPORTED -2 /* cannot be represented: pixel outside 0..255, foreign block grid */
#define ME_ERR_CUDA        -3 /* a CUDA call failed; see me_b200_last_error()                  */
#define ME_ERR_NO_DEVICE   -4 /* no usable CUDA device (there is no CPU fallback)              */
#define ME_ERR_NOMEM       -5
#define ME_ERR_STATE       -6 /* wait on an idle slot, submit on a busy one                    */

/* kernel selection (me_b200_create_ex).  AUTO picks the tuned kernel when the
 * geometry allows it (the small-span kernel for R <= 4) and the generic exact kernel otherwise;
 * all give identical results. */
#define ME_KERNEL_AUTO    0
#define ME_KERNEL_GENERIC 1
#define ME_KERNEL_TILED   2
#define ME_KERNEL_DIRECT  3 /* small spans (R <= 4): register-streaming kernels, the memory-bound end */
/* reported by me_b200_last_kernel only (not selectable: they follow the context's cost / search mode) */
#define ME_KERNEL_SSIM    4
#define ME_KERNEL_FAST    5

/* matching cost (me_b200_set_cost).  MSE: src/cpu/main.c:18-36.  SSIM: src/common/ssim.c:44-60,
 * maximised, as src/cpu/main_ssim.c runs it. */
#define ME_COST_MSE  0
#define ME_COST_SSIM 1

/* search pattern (me_b200_set_search).  FULL is the reference's exhaustive scan.  THREE_STEP and
 * DIAMOND are NOT in the reference (BASELINE.json config 4 names them): they are defined in
 * DESIGN.md section 5.6 / csrc/me_fast.cu on top of the reference's window (main.c:69-76), cost
 * (main.c:18-27) and tie rule (main.c:56) -- parity unpinned. */
#define ME_SEARCH_FULL       0
#define ME_SEARCH_THREE_STEP 1
#define ME_SEARCH_DIAMOND    2

#define ME_B200_MAX_SLOTS 4

typedef struct me_b200_ctx me_b200_ctx;

int         me_b200_abi_version(void);
const char *me_b200_strerror(int code);
/* text of the last CUDA failure seen by this context (or by the implicit
 * context of me_b200_search when ctx == NULL). */
const char *me_b200_last_error(const me_b200_ctx *ctx);
/* number of CUDA devices visible, 0 if none / no driver. */
int         me_b200_device_count(void);

/* ---- context: one per GPU and geometry --------------------------------------
 * replaces: the per-run setup of main.c:117-143 (blkDim, extraSpan, W, H;
 * createPredictionFrame).  Owns all device memory; caller owns all host memory. */
int  me_b200_create(me_b200_ctx **ctx, int device, int width, int height,
                    int blk_dim, int extra_span);
/* max_pairs: largest batch one call may carry (>=1); kernel: ME_KERNEL_*. */
int  me_b200_create_ex(me_b200_ctx **ctx, int device, int width, int height,
                       int blk_dim, int extra_span, int max_pairs, int kernel);
void me_b200_destroy(me_b200_ctx *ctx);

int      me_b200_num_blocks(const me_b200_ctx *ctx);   /* prediction_frame.c:9-12 */
int      me_b200_blocks_x(const me_b200_ctx *ctx);
int      me_b200_blocks_y(const me_b200_ctx *ctx);
/* the full-search MSE kernel this context runs: of the most recent launch once there has been one
 * (so a launch the tuned kernel could not take reports ME_KERNEL_GENERIC), before that the choice
 * made at create time for the context's own buffers.  ME_KERNEL_GENERIC, _TILED or _DIRECT. */
int      me_b200_kernel_in_use(const me_b200_ctx *ctx);
/* kernel family of the most recent search launch of any mode (ME_KERNEL_*, 0 = none yet) */
int      me_b200_last_kernel(const me_b200_ctx *ctx);
/* ME_KERNEL_AUTO searches that the tuned kernel could not serve and the generic kernel ran instead
 * (same results, ~10x slower); the reason of the last one is in me_b200_last_error(ctx).  0 in
 * every supported configuration -- tests and bench.py assert it. */
uint64_t me_b200_fallback_launches(const me_b200_ctx *ctx);
/* exact work counts of one frame pair */
uint64_t me_b200_pixel_compares(const me_b200_ctx *ctx);
uint64_t me_b200_candidates(const me_b200_ctx *ctx);
/* kernel launches issued by this context so far (search + pack kernels) */
uint64_t me_b200_launch_count(const me_b200_ctx *ctx);

/* ---- mode of a context (default: ME_COST_MSE, ME_SEARCH_FULL) ---------------------------
 * Every search entry point below that takes a ctx (search_u8, submit/wait, sequences, device,
 * band) then runs that cost / pattern with the same argument meaning.  With ME_COST_SSIM
 *   score[] = the float findBestBlkSSIM returns (main_ssim.c:29; 0 when no candidate scored > 0)
 *   ssd[]   = 1 when some candidate scored above 0, else 0.  In that case the reference leaves
 *             the motion vector uninitialised (ssim.c:88-103, main_ssim.c:26-27); here it is (0,0).
 * SSIM with a fast pattern is not defined (ME_ERR_UNSUPPORTED). */
int me_b200_set_cost(me_b200_ctx *ctx, int cost);
int me_b200_set_search(me_b200_ctx *ctx, int search);
/* candidate evaluations of all fast searches this context has run so far (synchronises). */
int me_b200_fast_evaluations(me_b200_ctx *ctx, uint64_t *evaluations);
/* first step of the three-step pattern: largest power of two <= max(1, (extra_span+1)/2). */
int me_b200_tss_first_step(int extra_span);

/* ---- reference drop-in ---------------------------------------------------------
 * replaces: the whole dispatch loop main.c:144-158, i.e. one call instead of
 * num_blks x thpool_add_work(runFindBestBlkMse).  pf->frame is the current
 * frame, refFrame the reference frame, both `int` per pixel (utils.c:49-53).
 * On success every pf->blks[i] has motion_vectorX/Y filled and
 * is_best_match_found = 1 (populateBlkMotionVector, main.c:11-15), which
 * motionCompensatedFrame requires (utils.c:105-108).
 * Optional out arrays (may be NULL), num_blks entries each: scores[i] = the
 * float findBestBlkMse returns (main.c:81), ssd[i] = exact integer SSD.
 * Uses an internal context cached per (device, W, H, B, R); device from env
 * ME_B200_DEVICE (default 0).  Blocking; not re-entrant for the same geometry. */
int me_b200_search(predictionFrame *pf, const int *refFrame, int extraSpan);
int me_b200_search_scores(predictionFrame *pf, const int *refFrame, int extraSpan,
                          float *scores, uint32_t *ssd);
/* SSIM twin -- replaces the sequential loop src/cpu/main_ssim.c:67-77 over findBestBlkSSIM
 * (main_ssim.c:16-30 -> ssim.c:83-107).  scores[i] = best SSIM, found[i] as ssd[] above. */
int me_b200_search_ssim(predictionFrame *pf, const int *refFrame, int extraSpan);
int me_b200_search_ssim_scores(predictionFrame *pf, const int *refFrame, int extraSpan,
                               float *scores, uint32_t *found);
/* fast-search twin on the reference's structs; search = ME_SEARCH_THREE_STEP / _DIAMOND. */
int me_b200_search_fast(predictionFrame *pf, const int *refFrame, int extraSpan, int search,
                        float *scores, uint32_t *ssd);
/* frees the cached implicit contexts (optional; also done at process exit). */
void me_b200_release_cached(void);

/* ---- host u8 path ----------------------------------------------------------------
 * replaces: yuvReadFrame's uint8 -> int widening (utils.c:61-73) + the search.
 * cur/ref: npairs frame pairs, 8-bit luma, row-major, stride == width, pair p at
 * cur + p*W*H.  Outputs: npairs*num_blocks entries each, any may be NULL.
 * Blocking: H2D, search, D2H. */
int me_b200_search_u8(me_b200_ctx *ctx, const uint8_t *cur, const uint8_t *ref, int npairs,
                      int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score);

/* pipelined form: up to ME_B200_MAX_SLOTS batches in flight, each slot on its
 * own stream (H2D of slot k+1 overlaps the search of slot k).  Host buffers
 * should be pinned (me_b200_host_alloc) for the copies to be asynchronous and
 * must stay valid until me_b200_wait(slot) returns. */
int me_b200_submit(me_b200_ctx *ctx, int slot, const uint8_t *cur, const uint8_t *ref, int npairs,
                   int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score);
int me_b200_wait(me_b200_ctx *ctx, int slot);
/* Ingest helper for boxes whose GPUs do not all have a full-speed host link (measured on this pool's 8-GPU
 * box, profiles/h2d_probe_r02.txt: four GPUs share one PCIe uplink and get 24 GB/s each under load, the
 * other four 36 GB/s; a 1080p +-32 search consumes 28.4 GB/s).  After this call the LAST helper_pairs
 * pairs of every me_b200_submit travel host -> helper_device (over THAT GPU's host link) -> this GPU
 * (cudaMemcpyPeerAsync over NVLink), the rest directly; results are unchanged.  helper_device: any other
 * GPU with peer access; helper_pairs < max_pairs; (-1, 0) switches it off.  Not during a submit in flight. */
int me_b200_set_ingest_helper(me_b200_ctx *ctx, int helper_device, int helper_pairs);
/* ---- sequences ----------------------------------------------
 * nframes consecutive frames of one video (u8, stride == width, frame i at frames + i*W*H),
 * 2 <= nframes <= max_pairs + 1.  Pair i searches frame i+1 (current) in frame i (reference),
 * i.e. the reference program run on every consecutive pair (argv[1] = frame i+1, argv[2] =
 * frame i, main.c:114-115).  Each frame crosses PCIe once and is used as the current frame of
 * one pair and the reference frame of the next.  Outputs: (nframes-1)*num_blocks entries.
 * me_b200_submit_sequence is asynchronous on the slot's stream (pair with me_b200_wait). */
int me_b200_submit_sequence(me_b200_ctx *ctx, int slot, const uint8_t *frames, int nframes,
                            int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score);
int me_b200_search_sequence_u8(me_b200_ctx *ctx, const uint8_t *frames, int nframes,
                               int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score);
void *me_b200_host_alloc(size_t bytes); /* pinned; NULL on failure */
/* flags: ME_HOST_WRITE_COMBINED = upload-only buffers (fast for the CPU to fill sequentially and for
 * the GPU to fetch, very slow for the CPU to read back) */
#define ME_HOST_WRITE_COMBINED 1
void *me_b200_host_alloc_ex(size_t bytes, int flags);
void  me_b200_host_free(void *p);

/* ---- device-resident path -----------------------------------------------------
 * All pointers are DEVICE pointers on ctx's device.  Frames: u8, row pitch
 * `pitch` bytes (>= width), pair p at d_cur + p*pair_stride.  Outputs
 * npairs*num_blocks entries, any may be NULL.  Enqueued on `stream` (a
 * cudaStream_t passed as void*, NULL = default stream); returns without
 * synchronising.  The tuned kernel needs 16-byte aligned base/pitch/pair_stride
 * (TMA); other layouts run the generic kernel. */
int me_b200_search_device(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                          size_t pitch, size_t pair_stride, int npairs,
                          int32_t *d_mvx, int32_t *d_mvy, uint32_t *d_ssd, float *d_score,
                          void *stream);
/* same, restricted to block rows [by_begin, by_end) of every pair -- the band
 * sharding of one very large frame over several GPUs.
 * Output arrays are still indexed by the global block index. */
int me_b200_search_device_band(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                               size_t pitch, size_t pair_stride, int npairs,
                               int by_begin, int by_end,
                               int32_t *d_mvx, int32_t *d_mvy, uint32_t *d_ssd, float *d_score,
                               void *stream);

/* ---- band sharding without a collective: peer-mapped fields over NVLink ---------------
 * One process per GPU.  Every rank allocates its copy of the field with me_b200_device_alloc,
 * exports it (CUDA IPC) and opens the copies of its peers; me_b200_search_device_band_peers
 * then searches the rank's block rows and stores every block's result into the local field AND
 * into each peer field -- from inside the search kernel where the tuned kernel runs, with a small
 * store kernel for rows another kernel produced.  me_b200_peer_barrier is a device-side barrier
 * between the GPUs (flags in the peer-mapped memory), enqueued on the same stream: when it has
 * passed on a rank, every peer's band has landed in that rank's field.  Replaces the
 * all_gather that followed me_b200_search_device_band. */
#define ME_B200_MAX_PEERS        8
#define ME_B200_IPC_HANDLE_BYTES 64
typedef struct me_b200_field {
  int32_t *mvx;
  int32_t *mvy;
  uint32_t *ssd;
  float *score; /* any member may be NULL */
} me_b200_field;
/* zero-filled device memory on ctx's device (cudaMalloc, so it can be exported); NULL on failure */
void *me_b200_device_alloc(me_b200_ctx *ctx, size_t bytes);
void  me_b200_device_free(me_b200_ctx *ctx, void *d_ptr);
int   me_b200_ipc_export(me_b200_ctx *ctx, void *d_ptr, unsigned char handle[ME_B200_IPC_HANDLE_BYTES]);
int   me_b200_ipc_open(me_b200_ctx *ctx, const unsigned char handle[ME_B200_IPC_HANDLE_BYTES], void **d_ptr);
int   me_b200_ipc_close(me_b200_ctx *ctx, void *d_ptr);
/* npeers = number of OTHER GPUs (0..ME_B200_MAX_PEERS-1); peers[i] = field of the i-th of them. */
int me_b200_search_device_band_peers(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                                     size_t pitch, size_t pair_stride, int npairs,
                                     int by_begin, int by_end,
                                     const me_b200_field *local, const me_b200_field *peers, int npeers,
                                     void *stream);
/* flags[q] = device pointer to rank q's flag array (>= nranks uint32, zero-initialised; flags[my_rank]
 * is this rank's own array, the others are peer-mapped).  epoch must grow by one per barrier.
 * Asynchronous on `stream`; gives up after timeout_ms (see me_b200_peer_barrier_timed_out). */
int me_b200_peer_barrier(me_b200_ctx *ctx, uint32_t *const *flags, int nranks, int my_rank,
                         uint32_t epoch, int timeout_ms, void *stream);
/* synchronises the device; *timed_out = 1 if any barrier of this context gave up since the last call */
int me_b200_peer_barrier_timed_out(me_b200_ctx *ctx, int *timed_out);

/* ---- post-search stage on device ------------------------
 * replaces: main.c:160-168 (motionCompensatedFrame + 2x frameDiff, utils.c:94-134).
 * d_out5: 5 stacked W x H u8 planes (ref, cur, mc, |ref-cur|, |mc-cur|), stride W.
 * d_sq_err (may be NULL): one uint64 = sum (mc-cur)^2, d_max (may be NULL): one
 * uint32 = max pixel of mc and cur -- the two inputs of imagePSNR (utils.c:137-164). */
int me_b200_postprocess_device(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                               size_t pitch, const int32_t *d_mvx, const int32_t *d_mvy,
                               uint8_t *d_out5, unsigned long long *d_sq_err, uint32_t *d_max,
                               void *stream);

/* batched form: pair p reads d_cur / d_ref + p*pair_stride and the p-th field (npairs*num_blocks entries,
 * as the search writes them), writes d_out5 + p*out_pair_stride (>= 5*W*H), d_sq_err[p] and d_max[p].
 * 16 pixels per thread (16-byte loads and stores) when W is a multiple of 16 and the buffers are
 * 16-byte aligned; HBM-bound: 2 bytes read + 5 written per pixel. */
int me_b200_postprocess_device_batch(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                                     size_t pitch, size_t pair_stride, int npairs,
                                     const int32_t *d_mvx, const int32_t *d_mvy,
                                     uint8_t *d_out5, size_t out_pair_stride,
                                     unsigned long long *d_sq_err, uint32_t *d_max, void *stream);

/* ---- integer-pipe microbenchmark (defines the roofline denominator) ---------------
 * Runs `which` (ME_PEAK_*) on `device` for about `iters` loop trips per thread
 * and returns lane-instructions per second of the named SASS op (0 on error). */
#define ME_PEAK_IDP4A        0 /* IDP.4A.U8.U8                          */
#define ME_PEAK_VABSDIFF4    1 /* VABSDIFF4.U8                          */
#define ME_PEAK_SSD_PAIR     2 /* VABSDIFF4 + IDP.4A pairs (4 px / pair) */
#define ME_PEAK_IADD3        3
#define ME_PEAK_LOP3         4
#define ME_PEAK_IMAD         5
#define ME_PEAK_VIMNMX       6
#define ME_PEAK_SSD_PAIR_LDS 7 /* pairs + 1 LDS.32 per 16 pairs          */
#define ME_PEAK_IDP4A_IADD3  8 /* IDP.4A + IADD3 1:1 (do the two pipes overlap?) */
#define ME_PEAK_LOOP_REPLICA 9 /* the search kernel's register pattern: IDP.4A lanes/s */
#define ME_PEAK_COUNT        10
double me_b200_int_peak(int device, int which, int iters, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif
