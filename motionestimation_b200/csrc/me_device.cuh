// me_device.cuh -- shared device-side declarations of the sm_100a search kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace me {

// Frame geometry + search parameters, passed by value to every kernel.
struct Geom {
  int W, H;        // frame size in pixels
  int B, R;        // block dimension, extra span (reference: blkDim, extraSpan)
  int nbx, nby;    // ceil(W/B), ceil(H/B)                (prediction_frame.c:9-10)
  int by_begin;    // first block row handled by this launch (band sharding)
  int by_count;    // number of block rows handled
};

// SoA outputs, indexed pair * (nbx*nby) + block index; any pointer may be null.
struct Out {
  int32_t *mvx;
  int32_t *mvy;
  uint32_t *ssd;
  float *score;
};

// Frames of a batch: u8, `pitch` bytes per row, pair p at base + p * pair_stride.
struct Frames {
  const uint8_t *cur;
  const uint8_t *ref;
  size_t pitch;
  size_t pair_stride;
};

// Stream-ordered scratch (energy tables, SSIM statistics tables, work counters): one memory pool
// per device, owned by this library and created on first use.  It keeps its memory between
// launches (release threshold = max; footprint = the largest scratch any launch needed, e.g.
// 4 B x W x rows per pair for the energy table, 8 B x W x rows per pair for the SSIM table) and
// never touches the attributes of the device's default pool, which the host application shares.
cudaError_t scratch_pool(cudaMemPool_t *pool);   // pool of the CURRENT device

// host-side launchers (defined in the .cu files)
cudaError_t launch_generic(const Geom &g, const Frames &f, int npairs, const Out &o, cudaStream_t s);

// Small-span kernel (R <= 4): one thread per (block, candidate); the memory-bound end of the path.
bool direct_supported(const Geom &g, size_t pitch, size_t pair_stride, const void *cur, const void *ref);
cudaError_t launch_direct(const Geom &g, const Frames &f, int npairs, const Out &o, cudaStream_t s);

// Tiled kernel: returns false from tiled_supported() when the geometry is not
// one it is specialised for (the caller then uses the generic kernel).
bool tiled_supported(const Geom &g, size_t pitch, size_t pair_stride, const void *cur, const void *ref);
struct TiledPlan;  // opaque: tensor maps + launch shape
cudaError_t tiled_plan_create(TiledPlan **plan, const Geom &g, int max_pairs);
void tiled_plan_destroy(TiledPlan *plan);
unsigned long long tiled_plan_launches(const TiledPlan *plan);  // kernels launched so far
// false after a failed launch_tiled that had not enqueued anything writing the caller's outputs yet
bool tiled_plan_outputs_enqueued(const TiledPlan *plan);
// Band sharding over NVLink: the next launch_tiled also stores every block's result into `peers`
// (output arrays in peer-mapped memory); tiled_plan_fused_rows reports the block rows it covered.
int tiled_plan_set_peers(TiledPlan *plan, const Out *peers, int npeers);
void tiled_plan_fused_rows(const TiledPlan *plan, int *begin, int *end);
cudaError_t launch_tiled(TiledPlan *plan, const Geom &g, const Frames &f, int npairs, const Out &o,
                         cudaStream_t s, const char **err_text);
// "Arriving frame": the next launch_tiled (whole frame, one pair) may start while the pair is still being
// uploaded top to bottom.  *flag (device memory) = base + number of current-frame rows resident, with the
// reference rows R below them; the uploader advances it in stream order after each piece.  Items wait for
// their rows; *status becomes 1 if one gave up after ~1 s.  (nullptr switches it off.)
// ref_resident: the WHOLE reference frame is already on the device (only the current frame is arriving) -- then the
// energy-table formulations (pre-pass over the reference frame + FORM 2 / 3) may serve the launch.
bool tiled_arrive_supported(const Geom &g);
void tiled_plan_set_arrive(TiledPlan *plan, const unsigned int *flag, unsigned int base, int *status,
                           bool ref_resident = false);

// SSIM cost on the tiled kernel (me_tiled.cu, FORM 4): full-height, full-width 16x16 block rows; the caller supplies
// the statistics tables ({pixel sum, stddev bits} per reference rectangle, row pitch W entries, and per current
// block of the launch).  cudaErrorInvalidConfiguration = geometry not covered, nothing was launched.
cudaError_t launch_tiled_ssim16(const Geom &g, const Frames &f, int npairs, const Out &o, int by_begin, int by_count,
                                const int2 *table, int table_y_lo, int table_rows, size_t table_pair_stride,
                                const int2 *blk_stats, cudaStream_t s, unsigned long long *launches);

// SSIM-cost full search (me_ssim.cu; reference: src/cpu/main_ssim.c + src/common/ssim.c).
// Out.score = best SSIM, Out.ssd = 1 when some candidate scored above 0 (else MV = (0,0)).
bool ssim_tiled_supported(const Geom &g, size_t pitch, size_t pair_stride, const void *cur, const void *ref);
cudaError_t launch_ssim(const Geom &g, const Frames &f, int npairs, const Out &o, bool tiled, cudaStream_t s,
                        unsigned long long *launches);

// Fast searches (me_fast.cu; absent from the reference, defined in that file's header).
// algo 1 = three-step, 2 = diamond; evals = optional device counter of candidate evaluations.
int tss_first_step(int R);
cudaError_t launch_fast(const Geom &g, const Frames &f, int npairs, const Out &o, int algo,
                        unsigned long long *evals, cudaStream_t s);

// Peer plumbing (me_peer.cu): copy block rows [by_begin, by_end) of a field into the peers'
// fields with plain stores over NVLink, and a device-side flag barrier between the GPUs.
cudaError_t launch_peer_scatter(const Geom &g, int npairs, int by_begin, int by_end, const Out &local,
                                const Out *peers, int npeers, cudaStream_t s);
cudaError_t launch_peer_barrier(uint32_t *const *flags, int npeers, int my_rank, uint32_t epoch,
                                unsigned long long timeout_ns, int *d_status, cudaStream_t s);

cudaError_t launch_postprocess(const Geom &g, const uint8_t *cur, const uint8_t *ref, size_t pitch,
                               const int32_t *mvx, const int32_t *mvy, uint8_t *out5,
                               unsigned long long *sq_err, uint32_t *mx, cudaStream_t s);

// batched: pair p reads cur/ref + p*pair_stride and the p-th field, writes out5 + p*out_pair_stride and
// sq_err[p] / mx[p]; 16 pixels per thread when the layout is 16-byte aligned, else per pixel
cudaError_t launch_postprocess_batch(const Geom &g, const uint8_t *cur, const uint8_t *ref, size_t pitch,
                                     size_t pair_stride, int npairs, const int32_t *mvx, const int32_t *mvy,
                                     uint8_t *out5, size_t out_pair_stride, unsigned long long *sq_err, uint32_t *mx,
                                     cudaStream_t s);

// order-preserving float -> u32 for non-negative floats
__device__ __forceinline__ uint32_t score_bits(float s) { return __float_as_uint(s); }

}  // namespace me
