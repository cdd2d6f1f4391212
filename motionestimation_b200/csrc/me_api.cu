// me_api.cu -- implementation of the C ABI declared in include/me_b200.h.
// Host-side plumbing only: contexts, device buffers, streams, copies, launches.
// All arithmetic of the search lives in me_generic.cu / me_tiled.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "me_b200.h"
#include "me_device.cuh"

struct me_slot {
  cudaStream_t stream = nullptr;
  uint8_t *d_cur = nullptr, *d_ref = nullptr;  // max_pairs frames each, pitch = ctx->pitch
  int32_t *d_mvx = nullptr, *d_mvy = nullptr;
  uint32_t *d_ssd = nullptr;
  float *d_score = nullptr;
  bool busy = false;
  // ingest helper (me_b200_set_ingest_helper): staging for the pairs that travel over the helper GPU's
  // host link, its stream and "upload landed" event on that device
  uint8_t *h_cur = nullptr, *h_ref = nullptr;
  cudaStream_t h_stream = nullptr;
  cudaEvent_t h_event = nullptr;
};

struct me_b200_ctx {
  int device = 0;
  me::Geom g{};
  int nb = 0;
  int max_pairs = 1;
  int kernel_req = ME_KERNEL_AUTO;
  int kernel = ME_KERNEL_GENERIC;  // what the internal (slot) path runs
  size_t pitch = 0, frame_bytes = 0;
  me_slot slots[ME_B200_MAX_SLOTS];
  me::TiledPlan *plan = nullptr;
  uint64_t launches = 0;
  uint64_t fallback_launches = 0;  // AUTO searches the tuned kernel could not serve (generic kernel ran)
  int helper_device = -1;          // ingest helper: another GPU whose host link carries part of every upload
  int helper_pairs = 0;            // ... the last helper_pairs pairs of a submit
  int last_kernel = 0;             // kernel of the most recent search launch (0: none yet)
  int cost = ME_COST_MSE;          // me_b200_set_cost
  int search = ME_SEARCH_FULL;     // me_b200_set_search
  unsigned long long *d_evals = nullptr;  // fast search: candidate evaluations so far
  int *d_peer_status = nullptr;           // peer barrier: 1 after a time-out
  cudaEvent_t band_events[32] = {nullptr}; // drop-in call: "band c has been uploaded" (created on first use)
  // drop-in call, arriving-frame mode: device flag "rows resident" + give-up status, the pinned values the
  // copy stream writes into the flag, and the call counter that keeps flag values of different calls apart
  unsigned int *d_arrive = nullptr;        // [0] flag, [1] status
  unsigned int *h_arrive = nullptr;        // pinned, one value per band
  unsigned int arrive_epoch = 0;
  char err[256] = {0};
  // scratch for the int-frame drop-in path
  uint8_t *h_cur = nullptr, *h_ref = nullptr;  // pinned, W*H each
  int32_t *h_mvx = nullptr, *h_mvy = nullptr;
  uint32_t *h_ssd = nullptr;
  float *h_score = nullptr;
};

namespace {

char g_err[256] = {0};

// one library-owned stream-ordered memory pool per device (see me_device.cuh)
std::mutex g_pool_mu;
cudaMemPool_t g_pools[64] = {nullptr};
}  // namespace

cudaError_t me::scratch_pool(cudaMemPool_t *pool) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pools[dev]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    e = cudaMemPoolCreate(&g_pools[dev], &props);
    if (e != cudaSuccess) {
      g_pools[dev] = nullptr;
      return e;
    }
    unsigned long long keep = ~0ull;   // keep the scratch between launches
    cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    (void)cudaGetLastError();
  }
  *pool = g_pools[dev];
  return cudaSuccess;
}

namespace {

int fail_cuda(me_b200_ctx *ctx, cudaError_t e, const char *what) {
  char *dst = ctx ? ctx->err : g_err;
  snprintf(dst, 256, "%s: %s", what, cudaGetErrorString(e));
  (void)cudaGetLastError();
  return ME_ERR_CUDA;
}

#define ME_CUDA(ctx, call)                                    \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call); \
  } while (0)

uint64_t axis_sum(int N, int B, int R, bool weighted) {
  uint64_t s = 0;
  for (int p = 0; p < N; p += B) {
    int e = p + B < N ? B : N - p;
    int lo = p - R < 0 ? 0 : p - R;
    int hi = p + e - 1 + R >= N ? N - 1 : p + e - 1 + R;
    uint64_t nc = (uint64_t)(hi - e + 1 - lo + 1);
    s += weighted ? nc * (uint64_t)e : nc;
  }
  return s;
}


// Pick + launch the search for frames already on the device.
int run_search(me_b200_ctx *ctx, const me::Frames &f, int npairs, int by_begin, int by_end,
               const me::Out &o, cudaStream_t s) {
  me::Geom g = ctx->g;
  g.by_begin = by_begin;
  g.by_count = by_end - by_begin;
  if (g.by_count <= 0 || npairs <= 0) return ME_OK;
  if (ctx->search != ME_SEARCH_FULL) {
    // three-step / diamond search (MSE cost), one warp per block
    cudaError_t e = me::launch_fast(g, f, npairs, o, ctx->search, ctx->d_evals, s);
    if (e == cudaErrorInvalidConfiguration) {
      (void)cudaGetLastError();
      snprintf(ctx->err, 256, "fast search: block size %d does not fit shared memory", g.B);
      return ME_ERR_UNSUPPORTED;
    }
    if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_fast");
    ctx->launches += (uint64_t)((npairs + 65534) / 65535);
    ctx->last_kernel = ME_KERNEL_FAST;
    return ME_OK;
  }
  if (ctx->cost == ME_COST_SSIM) {
    unsigned long long n = 0;
    cudaError_t e = me::launch_ssim(g, f, npairs, o, ctx->kernel_req != ME_KERNEL_GENERIC, s, &n);
    ctx->launches += n;
    if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_ssim");
    ctx->last_kernel = ME_KERNEL_SSIM;
    return ME_OK;
  }
  // small spans: one thread per (block, candidate)
  const bool direct_ok = me::direct_supported(g, f.pitch, f.pair_stride, f.cur, f.ref);
  if (ctx->kernel_req == ME_KERNEL_DIRECT && !direct_ok) {
    snprintf(ctx->err, 256, "small-span kernel requested but geometry/layout unsupported");
    return ME_ERR_UNSUPPORTED;
  }
  if (direct_ok && (ctx->kernel_req == ME_KERNEL_AUTO || ctx->kernel_req == ME_KERNEL_DIRECT)) {
    cudaError_t e = me::launch_direct(g, f, npairs, o, s);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_direct");
    ctx->launches += (uint64_t)((npairs + 65534) / 65535);
    ctx->last_kernel = ME_KERNEL_DIRECT;
    return ME_OK;
  }
  bool tiled = ctx->kernel_req != ME_KERNEL_GENERIC && ctx->plan &&
               me::tiled_supported(g, f.pitch, f.pair_stride, f.cur, f.ref);
  if (ctx->kernel_req == ME_KERNEL_TILED && !tiled) {
    snprintf(ctx->err, 256, "tiled kernel requested but geometry/layout unsupported");
    return ME_ERR_UNSUPPORTED;
  }
  if (tiled) {
    const char *txt = nullptr;
    cudaError_t e = me::launch_tiled(ctx->plan, g, f, npairs, o, s, &txt);
    if (e != cudaSuccess && ctx->kernel_req != ME_KERNEL_TILED && !me::tiled_plan_outputs_enqueued(ctx->plan)) {
      // AUTO: the tuned kernel could not take this launch (window does not fit its shared-memory
      // ring, tensor-map encode rejected the layout, scratch allocation failed ...) and nothing that
      // writes the outputs has been enqueued: the generic kernel serves it with identical results.
      // Never silent: counted in me_b200_fallback_launches, reason kept in me_b200_last_error,
      // and me_b200_kernel_in_use reports ME_KERNEL_GENERIC from now on.
      (void)cudaGetLastError();
      snprintf(ctx->err, 256, "tuned kernel not used (%s: %s); generic kernel ran", txt ? txt : "launch_tiled",
               cudaGetErrorString(e));
      ctx->fallback_launches++;
      tiled = false;
    } else if (e != cudaSuccess) {
      return fail_cuda(ctx, e, txt ? txt : "launch_tiled");
    }
  }
  if (tiled) {
    ctx->last_kernel = ME_KERNEL_TILED;
    return ME_OK;  // counted by the plan (search kernel + pre-pass kernels)
  }
  ctx->last_kernel = ME_KERNEL_GENERIC;
  {
    // generic kernel; grid.y carries the pair index
    int done = 0;
    while (done < npairs) {
      int n = npairs - done > 65535 ? 65535 : npairs - done;
      me::Frames ff = f;
      ff.cur += (size_t)done * f.pair_stride;
      ff.ref += (size_t)done * f.pair_stride;
      me::Out oo = o;
      size_t off = (size_t)done * ctx->nb;
      if (oo.mvx) oo.mvx += off;
      if (oo.mvy) oo.mvy += off;
      if (oo.ssd) oo.ssd += off;
      if (oo.score) oo.score += off;
      cudaError_t e = me::launch_generic(g, ff, n, oo, s);
      if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_generic");
      ctx->launches++;
      done += n;
    }
  }
  return ME_OK;
}

int use_device(me_b200_ctx *ctx) {
  ME_CUDA(ctx, cudaSetDevice(ctx->device));
  return ME_OK;
}

// Second half of a submit, after the uploads have been queued on the slot's stream: search + the
// downloads of the requested fields.  On a failure the stream is drained before returning, so the
// caller's host buffers are no longer read or written once the error is reported.
int finish_submit(me_b200_ctx *ctx, me_slot &s, const me::Frames &f, int npairs, int32_t *mvx, int32_t *mvy,
                  uint32_t *ssd, float *score) {
  me::Out o{s.d_mvx, s.d_mvy, s.d_ssd, s.d_score};
  int rc = run_search(ctx, f, npairs, 0, ctx->g.nby, o, s.stream);
  const size_t ob = (size_t)ctx->nb * (size_t)npairs * 4;
  cudaError_t e = cudaSuccess;
  if (rc == ME_OK && mvx) e = cudaMemcpyAsync(mvx, s.d_mvx, ob, cudaMemcpyDeviceToHost, s.stream);
  if (rc == ME_OK && e == cudaSuccess && mvy) e = cudaMemcpyAsync(mvy, s.d_mvy, ob, cudaMemcpyDeviceToHost, s.stream);
  if (rc == ME_OK && e == cudaSuccess && ssd) e = cudaMemcpyAsync(ssd, s.d_ssd, ob, cudaMemcpyDeviceToHost, s.stream);
  if (rc == ME_OK && e == cudaSuccess && score)
    e = cudaMemcpyAsync(score, s.d_score, ob, cudaMemcpyDeviceToHost, s.stream);
  if (rc == ME_OK && e != cudaSuccess) rc = fail_cuda(ctx, e, "cudaMemcpyAsync(device to host)");
  if (rc != ME_OK) {
    cudaStreamSynchronize(s.stream);
    (void)cudaGetLastError();
    return rc;
  }
  s.busy = true;
  return ME_OK;
}

}  // namespace

extern "C" {

int me_b200_abi_version(void) { return ME_B200_ABI_VERSION; }

const char *me_b200_strerror(int code) {
  switch (code) {
    case ME_OK: return "ok";
    case ME_ERR_INVALID_ARG: return "invalid argument";
    case ME_ERR_UNSUPPORTED: return "unsupported input (not representable)";
    case ME_ERR_CUDA: return "CUDA error";
    case ME_ERR_NO_DEVICE: return "no usable CUDA device (no CPU fallback exists)";
    case ME_ERR_NOMEM: return "out of memory";
    case ME_ERR_STATE: return "slot state error";
    default: return "unknown error";
  }
}

const char *me_b200_last_error(const me_b200_ctx *ctx) { return ctx ? ctx->err : g_err; }

int me_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int me_b200_create(me_b200_ctx **ctx, int device, int width, int height, int blk_dim, int extra_span) {
  return me_b200_create_ex(ctx, device, width, height, blk_dim, extra_span, 1, ME_KERNEL_AUTO);
}

int me_b200_create_ex(me_b200_ctx **out, int device, int width, int height, int blk_dim,
                      int extra_span, int max_pairs, int kernel) {
  if (!out) return ME_ERR_INVALID_ARG;
  *out = nullptr;
  if (width <= 0 || height <= 0 || blk_dim <= 0 || extra_span < 0 || max_pairs < 1) return ME_ERR_INVALID_ARG;
  if (kernel != ME_KERNEL_AUTO && kernel != ME_KERNEL_GENERIC && kernel != ME_KERNEL_TILED &&
      kernel != ME_KERNEL_DIRECT)
    return ME_ERR_INVALID_ARG;
  if (blk_dim > 256 || extra_span > 1024) return ME_ERR_UNSUPPORTED;
  if ((long long)width * height > (1ll << 30)) return ME_ERR_UNSUPPORTED;
  int n = me_b200_device_count();
  if (n <= 0 || device < 0 || device >= n) return ME_ERR_NO_DEVICE;

  me_b200_ctx *ctx = new (std::nothrow) me_b200_ctx();
  if (!ctx) return ME_ERR_NOMEM;
  ctx->device = device;
  ctx->g.W = width;
  ctx->g.H = height;
  ctx->g.B = blk_dim;
  ctx->g.R = extra_span;
  ctx->g.nbx = (width + blk_dim - 1) / blk_dim;
  ctx->g.nby = (height + blk_dim - 1) / blk_dim;
  ctx->g.by_begin = 0;
  ctx->g.by_count = ctx->g.nby;
  ctx->nb = ctx->g.nbx * ctx->g.nby;
  ctx->max_pairs = max_pairs;
  ctx->kernel_req = kernel;
  ctx->pitch = ((size_t)width + 15) & ~(size_t)15;  // TMA: global row stride multiple of 16 B
  ctx->frame_bytes = ctx->pitch * (size_t)height;

  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    fail_cuda(nullptr, e, "cudaSetDevice");
    delete ctx;
    return ME_ERR_NO_DEVICE;
  }
  int rc = ME_OK;
  auto check = [&](cudaError_t ce, const char *what) {
    if (ce != cudaSuccess && rc == ME_OK) {
      fail_cuda(nullptr, ce, what);
      rc = ce == cudaErrorMemoryAllocation ? ME_ERR_NOMEM : ME_ERR_CUDA;
    }
  };
  const size_t fb = ctx->frame_bytes * ((size_t)max_pairs + 1) + 256;  // +1 frame: sequences; tail slack
  const size_t ob = (size_t)ctx->nb * (size_t)max_pairs;
  for (int i = 0; i < ME_B200_MAX_SLOTS && rc == ME_OK; i++) {
    me_slot &s = ctx->slots[i];
    check(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    check(cudaMalloc(&s.d_cur, fb), "cudaMalloc(cur)");
    check(cudaMalloc(&s.d_ref, fb), "cudaMalloc(ref)");
    check(cudaMalloc(&s.d_mvx, ob * 4), "cudaMalloc(mvx)");
    check(cudaMalloc(&s.d_mvy, ob * 4), "cudaMalloc(mvy)");
    check(cudaMalloc(&s.d_ssd, ob * 4), "cudaMalloc(ssd)");
    check(cudaMalloc(&s.d_score, ob * 4), "cudaMalloc(score)");
    if (rc == ME_OK) {
      check(cudaMemsetAsync(s.d_cur, 0, fb, s.stream), "memset");
      check(cudaMemsetAsync(s.d_ref, 0, fb, s.stream), "memset");
    }
  }
  if (rc == ME_OK && kernel != ME_KERNEL_GENERIC && kernel != ME_KERNEL_DIRECT) {
    cudaError_t pe = me::tiled_plan_create(&ctx->plan, ctx->g, max_pairs);
    if (pe != cudaSuccess) {
      ctx->plan = nullptr;
      (void)cudaGetLastError();
      if (kernel == ME_KERNEL_TILED) {
        fail_cuda(nullptr, pe, "tiled_plan_create");
        rc = ME_ERR_UNSUPPORTED;
      }
    }
  }
  if (rc == ME_OK) {
    me::Frames f{ctx->slots[0].d_cur, ctx->slots[0].d_ref, ctx->pitch, ctx->frame_bytes};
    ctx->kernel = (ctx->plan && me::tiled_supported(ctx->g, f.pitch, f.pair_stride, f.cur, f.ref))
                      ? ME_KERNEL_TILED
                      : ME_KERNEL_GENERIC;
    if ((kernel == ME_KERNEL_AUTO || kernel == ME_KERNEL_DIRECT) &&
        me::direct_supported(ctx->g, f.pitch, f.pair_stride, f.cur, f.ref))
      ctx->kernel = ME_KERNEL_DIRECT;
    if (kernel == ME_KERNEL_DIRECT && ctx->kernel != ME_KERNEL_DIRECT) {
      snprintf(g_err, 256, "small-span kernel does not support B=%d R=%d", blk_dim, extra_span);
      rc = ME_ERR_UNSUPPORTED;
    }
    if (kernel == ME_KERNEL_TILED && ctx->kernel != ME_KERNEL_TILED) {
      snprintf(g_err, 256, "tiled kernel does not support B=%d R=%d", blk_dim, extra_span);
      rc = ME_ERR_UNSUPPORTED;
    }
  }
  if (rc == ME_OK) check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  if (rc != ME_OK) {
    me_b200_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return ME_OK;
}

void me_b200_destroy(me_b200_ctx *ctx) {
  if (!ctx) return;
  if (cudaSetDevice(ctx->device) == cudaSuccess) {
    for (int i = 0; i < ME_B200_MAX_SLOTS; i++) {
      me_slot &s = ctx->slots[i];
      if (s.stream) cudaStreamSynchronize(s.stream);
      if (s.h_stream) {
        cudaStreamSynchronize(s.h_stream);
        cudaStreamDestroy(s.h_stream);
      }
      if (s.h_event) cudaEventDestroy(s.h_event);
      cudaFree(s.h_cur);   // (helper-device memory; cudaFree takes any device's pointer)
      cudaFree(s.h_ref);
      cudaFree(s.d_cur);
      cudaFree(s.d_ref);
      cudaFree(s.d_mvx);
      cudaFree(s.d_mvy);
      cudaFree(s.d_ssd);
      cudaFree(s.d_score);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (ctx->plan) me::tiled_plan_destroy(ctx->plan);
    cudaFree(ctx->d_evals);
    cudaFree(ctx->d_peer_status);
    cudaFree(ctx->d_arrive);
    cudaFreeHost(ctx->h_arrive);
    for (int i = 0; i < 32; i++)
      if (ctx->band_events[i]) cudaEventDestroy(ctx->band_events[i]);
    cudaFreeHost(ctx->h_cur);
    cudaFreeHost(ctx->h_ref);
    cudaFreeHost(ctx->h_mvx);
    cudaFreeHost(ctx->h_mvy);
    cudaFreeHost(ctx->h_ssd);
    cudaFreeHost(ctx->h_score);
    (void)cudaGetLastError();
  }
  delete ctx;
}

int me_b200_num_blocks(const me_b200_ctx *ctx) { return ctx ? ctx->nb : 0; }
int me_b200_blocks_x(const me_b200_ctx *ctx) { return ctx ? ctx->g.nbx : 0; }
int me_b200_blocks_y(const me_b200_ctx *ctx) { return ctx ? ctx->g.nby : 0; }
int me_b200_kernel_in_use(const me_b200_ctx *ctx) {
  if (!ctx) return 0;
  // the kernel of the most recent full-search MSE launch; before the first launch, the choice made
  // for the context's own buffers at create time
  const int k = ctx->last_kernel;
  if (k == ME_KERNEL_GENERIC || k == ME_KERNEL_TILED || k == ME_KERNEL_DIRECT) return k;
  return ctx->kernel;
}
int me_b200_last_kernel(const me_b200_ctx *ctx) { return ctx ? ctx->last_kernel : 0; }
uint64_t me_b200_fallback_launches(const me_b200_ctx *ctx) { return ctx ? ctx->fallback_launches : 0; }
uint64_t me_b200_pixel_compares(const me_b200_ctx *ctx) {
  if (!ctx) return 0;
  return axis_sum(ctx->g.W, ctx->g.B, ctx->g.R, true) * axis_sum(ctx->g.H, ctx->g.B, ctx->g.R, true);
}
uint64_t me_b200_candidates(const me_b200_ctx *ctx) {
  if (!ctx) return 0;
  return axis_sum(ctx->g.W, ctx->g.B, ctx->g.R, false) * axis_sum(ctx->g.H, ctx->g.B, ctx->g.R, false);
}
uint64_t me_b200_launch_count(const me_b200_ctx *ctx) {
  return ctx ? ctx->launches + me::tiled_plan_launches(ctx->plan) : 0;
}

int me_b200_set_cost(me_b200_ctx *ctx, int cost) {
  if (!ctx || (cost != ME_COST_MSE && cost != ME_COST_SSIM)) return ME_ERR_INVALID_ARG;
  if (cost == ME_COST_SSIM && ctx->search != ME_SEARCH_FULL) return ME_ERR_UNSUPPORTED;
  ctx->cost = cost;
  return ME_OK;
}

int me_b200_set_search(me_b200_ctx *ctx, int search) {
  if (!ctx || (search != ME_SEARCH_FULL && search != ME_SEARCH_THREE_STEP && search != ME_SEARCH_DIAMOND))
    return ME_ERR_INVALID_ARG;
  if (search != ME_SEARCH_FULL && ctx->cost != ME_COST_MSE) return ME_ERR_UNSUPPORTED;
  if (search != ME_SEARCH_FULL && !ctx->d_evals) {
    int rc = use_device(ctx);
    if (rc) return rc;
    ME_CUDA(ctx, cudaMalloc((void **)&ctx->d_evals, 256));
    ME_CUDA(ctx, cudaMemset(ctx->d_evals, 0, 256));
  }
  ctx->search = search;
  return ME_OK;
}

int me_b200_fast_evaluations(me_b200_ctx *ctx, uint64_t *evaluations) {
  if (!ctx || !evaluations) return ME_ERR_INVALID_ARG;
  *evaluations = 0;
  if (!ctx->d_evals) return ME_OK;
  int rc = use_device(ctx);
  if (rc) return rc;
  ME_CUDA(ctx, cudaDeviceSynchronize());
  unsigned long long v = 0;
  ME_CUDA(ctx, cudaMemcpy(&v, ctx->d_evals, sizeof v, cudaMemcpyDeviceToHost));
  *evaluations = v;
  return ME_OK;
}

int me_b200_tss_first_step(int extra_span) { return me::tss_first_step(extra_span); }

void *me_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}
void *me_b200_host_alloc_ex(size_t bytes, int flags) {
  void *p = nullptr;
  const unsigned f = (flags & ME_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, f) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}
void me_b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int me_b200_submit(me_b200_ctx *ctx, int slot, const uint8_t *cur, const uint8_t *ref, int npairs,
                   int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score) {
  if (!ctx || !cur || !ref || slot < 0 || slot >= ME_B200_MAX_SLOTS) return ME_ERR_INVALID_ARG;
  if (npairs < 1 || npairs > ctx->max_pairs) return ME_ERR_INVALID_ARG;
  me_slot &s = ctx->slots[slot];
  if (s.busy) return ME_ERR_STATE;
  int rc = use_device(ctx);
  if (rc) return rc;
  const size_t W = (size_t)ctx->g.W, H = (size_t)ctx->g.H;
  // pairs are contiguous on the host (stride W) and on the device (pitch): one copy per frame
  // set -- linear when the device pitch equals the width (no per-row DMA descriptors), 2-D otherwise
  cudaError_t e = cudaSuccess;
  // ingest helper: the last `hp` pairs go host -> helper GPU (its PCIe link) -> this GPU (NVLink peer copy)
  int hp = ctx->helper_device >= 0 ? ctx->helper_pairs : 0;
  if (hp > npairs - 1) hp = npairs - 1;
  if (hp < 0) hp = 0;
  const size_t nd = (size_t)(npairs - hp);
  auto h2d = [&](uint8_t *dst, const uint8_t *src, size_t frames, cudaStream_t st) {
    if (ctx->pitch == W) return cudaMemcpyAsync(dst, src, W * H * frames, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, ctx->pitch, src, W, W, H * frames, cudaMemcpyHostToDevice, st);
  };
  if (hp > 0) {
    // issued first: the detour has one more hop than the direct upload
    e = h2d(s.h_cur, cur + nd * W * H, (size_t)hp, s.h_stream);
    if (e == cudaSuccess) e = h2d(s.h_ref, ref + nd * W * H, (size_t)hp, s.h_stream);
    if (e == cudaSuccess) e = cudaEventRecord(s.h_event, s.h_stream);
  }
  if (e == cudaSuccess) e = h2d(s.d_cur, cur, nd, s.stream);
  if (e == cudaSuccess) e = h2d(s.d_ref, ref, nd, s.stream);
  if (e == cudaSuccess && hp > 0) {
    e = cudaStreamWaitEvent(s.stream, s.h_event, 0);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(s.d_cur + nd * ctx->frame_bytes, ctx->device, s.h_cur, ctx->helper_device,
                              (size_t)hp * ctx->frame_bytes, s.stream);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(s.d_ref + nd * ctx->frame_bytes, ctx->device, s.h_ref, ctx->helper_device,
                              (size_t)hp * ctx->frame_bytes, s.stream);
  }
  if (e != cudaSuccess) {
    if (s.h_stream) cudaStreamSynchronize(s.h_stream);
    cudaStreamSynchronize(s.stream);  // the first upload may still be reading the caller's buffer
    return fail_cuda(ctx, e, "cudaMemcpyAsync(host to device)");
  }
  me::Frames f{s.d_cur, s.d_ref, ctx->pitch, ctx->frame_bytes};
  return finish_submit(ctx, s, f, npairs, mvx, mvy, ssd, score);
}

int me_b200_set_ingest_helper(me_b200_ctx *ctx, int helper_device, int helper_pairs) {
  if (!ctx) return ME_ERR_INVALID_ARG;
  for (int i = 0; i < ME_B200_MAX_SLOTS; i++)
    if (ctx->slots[i].busy) return ME_ERR_STATE;
  if (helper_device < 0 || helper_pairs <= 0) {   // switch it off (the staging stays allocated until destroy)
    ctx->helper_device = -1;
    ctx->helper_pairs = 0;
    return ME_OK;
  }
  if (helper_device == ctx->device || helper_device >= me_b200_device_count()) return ME_ERR_INVALID_ARG;
  if (helper_pairs >= ctx->max_pairs) return ME_ERR_INVALID_ARG;
  if (ctx->helper_device >= 0 && ctx->helper_device != helper_device) return ME_ERR_UNSUPPORTED;   // one helper per context
  int a = 0, b = 0;
  ME_CUDA(ctx, cudaDeviceCanAccessPeer(&a, ctx->device, helper_device));
  ME_CUDA(ctx, cudaDeviceCanAccessPeer(&b, helper_device, ctx->device));
  if (!a || !b) {
    snprintf(ctx->err, 256, "devices %d and %d cannot access each other's memory", ctx->device, helper_device);
    return ME_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e == cudaSuccess) {
    e = cudaDeviceEnablePeerAccess(helper_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
  }
  if (e == cudaSuccess) e = cudaSetDevice(helper_device);
  if (e == cudaSuccess) {
    e = cudaDeviceEnablePeerAccess(ctx->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
  }
  const size_t bytes = ctx->frame_bytes * (size_t)helper_pairs + 256;
  for (int i = 0; i < ME_B200_MAX_SLOTS && e == cudaSuccess; i++) {
    me_slot &s = ctx->slots[i];
    if (s.h_cur && helper_pairs > ctx->helper_pairs) {   // grow
      cudaFree(s.h_cur); cudaFree(s.h_ref);
      s.h_cur = s.h_ref = nullptr;
    }
    if (!s.h_cur) {
      e = cudaMalloc(&s.h_cur, bytes);
      if (e == cudaSuccess) e = cudaMalloc(&s.h_ref, bytes);
    }
    if (e == cudaSuccess && !s.h_stream) e = cudaStreamCreateWithFlags(&s.h_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && !s.h_event) e = cudaEventCreateWithFlags(&s.h_event, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "me_b200_set_ingest_helper");
  ctx->helper_device = helper_device;
  ctx->helper_pairs = helper_pairs;
  return ME_OK;
}

int me_b200_submit_sequence(me_b200_ctx *ctx, int slot, const uint8_t *frames, int nframes, int32_t *mvx,
                            int32_t *mvy, uint32_t *ssd, float *score) {
  if (!ctx || !frames || slot < 0 || slot >= ME_B200_MAX_SLOTS) return ME_ERR_INVALID_ARG;
  if (nframes < 2 || nframes - 1 > ctx->max_pairs) return ME_ERR_INVALID_ARG;
  me_slot &s = ctx->slots[slot];
  if (s.busy) return ME_ERR_STATE;
  int rc = use_device(ctx);
  if (rc) return rc;
  const size_t W = (size_t)ctx->g.W, H = (size_t)ctx->g.H;
  // every frame is uploaded once into the slot's frame buffer
  if (ctx->pitch == W) {
    ME_CUDA(ctx, cudaMemcpyAsync(s.d_cur, frames, W * H * (size_t)nframes, cudaMemcpyHostToDevice, s.stream));
  } else {
    ME_CUDA(ctx, cudaMemcpy2DAsync(s.d_cur, ctx->pitch, frames, W, W, H * (size_t)nframes,
                                   cudaMemcpyHostToDevice, s.stream));
  }
  // pair i: current = frame i+1, reference = frame i -- two views of the same buffer
  const int npairs = nframes - 1;
  me::Frames f{s.d_cur + ctx->frame_bytes, s.d_cur, ctx->pitch, ctx->frame_bytes};
  return finish_submit(ctx, s, f, npairs, mvx, mvy, ssd, score);
}

int me_b200_search_sequence_u8(me_b200_ctx *ctx, const uint8_t *frames, int nframes, int32_t *mvx,
                               int32_t *mvy, uint32_t *ssd, float *score) {
  if (!ctx || !frames || nframes < 2) return ME_ERR_INVALID_ARG;
  // longer sequences go through in chunks of max_pairs pairs; consecutive chunks share one frame
  const size_t fsz = (size_t)ctx->g.W * ctx->g.H;
  int first = 0, k = 0, rc = ME_OK;
  int inflight[ME_B200_MAX_SLOTS] = {0, 0, 0, 0};
  while (first < nframes - 1 && rc == ME_OK) {
    const int slot = k % ME_B200_MAX_SLOTS;
    if (inflight[slot]) {
      rc = me_b200_wait(ctx, slot);
      inflight[slot] = 0;
      if (rc) break;
    }
    int np = nframes - 1 - first;
    if (np > ctx->max_pairs) np = ctx->max_pairs;
    const size_t oo = (size_t)first * ctx->nb;
    rc = me_b200_submit_sequence(ctx, slot, frames + (size_t)first * fsz, np + 1, mvx ? mvx + oo : nullptr,
                                 mvy ? mvy + oo : nullptr, ssd ? ssd + oo : nullptr, score ? score + oo : nullptr);
    if (rc == ME_OK) inflight[slot] = 1;
    first += np;
    k++;
  }
  for (int i = 0; i < ME_B200_MAX_SLOTS; i++)
    if (inflight[i]) {
      int r2 = me_b200_wait(ctx, i);
      if (rc == ME_OK) rc = r2;
    }
  return rc;
}

int me_b200_wait(me_b200_ctx *ctx, int slot) {
  if (!ctx || slot < 0 || slot >= ME_B200_MAX_SLOTS) return ME_ERR_INVALID_ARG;
  me_slot &s = ctx->slots[slot];
  if (!s.busy) return ME_ERR_STATE;
  s.busy = false;
  ME_CUDA(ctx, cudaStreamSynchronize(s.stream));
  return ME_OK;
}

int me_b200_search_u8(me_b200_ctx *ctx, const uint8_t *cur, const uint8_t *ref, int npairs,
                      int32_t *mvx, int32_t *mvy, uint32_t *ssd, float *score) {
  if (!ctx || !cur || !ref || npairs < 1) return ME_ERR_INVALID_ARG;
  // batches larger than the context was sized for go through in max_pairs chunks
  const size_t fsz = (size_t)ctx->g.W * ctx->g.H;
  int done = 0, inflight[ME_B200_MAX_SLOTS], k = 0;
  for (int i = 0; i < ME_B200_MAX_SLOTS; i++) inflight[i] = 0;
  int rc = ME_OK;
  while (done < npairs && rc == ME_OK) {
    int slot = k % ME_B200_MAX_SLOTS;
    if (inflight[slot]) {
      rc = me_b200_wait(ctx, slot);
      inflight[slot] = 0;
      if (rc) break;
    }
    int n = npairs - done > ctx->max_pairs ? ctx->max_pairs : npairs - done;
    size_t oo = (size_t)done * ctx->nb;
    rc = me_b200_submit(ctx, slot, cur + done * fsz, ref + done * fsz, n, mvx ? mvx + oo : nullptr,
                        mvy ? mvy + oo : nullptr, ssd ? ssd + oo : nullptr, score ? score + oo : nullptr);
    if (rc == ME_OK) inflight[slot] = 1;
    done += n;
    k++;
  }
  for (int i = 0; i < ME_B200_MAX_SLOTS; i++)
    if (inflight[i]) {
      int r2 = me_b200_wait(ctx, i);
      if (rc == ME_OK) rc = r2;
    }
  return rc;
}

int me_b200_search_device_band(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                               size_t pitch, size_t pair_stride, int npairs, int by_begin, int by_end,
                               int32_t *d_mvx, int32_t *d_mvy, uint32_t *d_ssd, float *d_score,
                               void *stream) {
  if (!ctx || !d_cur || !d_ref || npairs < 1) return ME_ERR_INVALID_ARG;
  if (pitch < (size_t)ctx->g.W) return ME_ERR_INVALID_ARG;
  if (npairs > 1 && pair_stride < pitch * (size_t)ctx->g.H) return ME_ERR_INVALID_ARG;
  if (by_begin < 0 || by_end > ctx->g.nby || by_begin > by_end) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  me::Frames f{d_cur, d_ref, pitch, pair_stride};
  me::Out o{d_mvx, d_mvy, d_ssd, d_score};
  return run_search(ctx, f, npairs, by_begin, by_end, o, (cudaStream_t)stream);
}

// ---- band sharding over peer-mapped memory ------------------------------------------------

void *me_b200_device_alloc(me_b200_ctx *ctx, size_t bytes) {
  if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  void *p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess || cudaMemset(p, 0, bytes ? bytes : 1) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return nullptr;
  }
  cudaDeviceSynchronize();
  return p;
}

void me_b200_device_free(me_b200_ctx *ctx, void *d_ptr) {
  if (ctx && d_ptr && cudaSetDevice(ctx->device) == cudaSuccess) cudaFree(d_ptr);
}

int me_b200_ipc_export(me_b200_ctx *ctx, void *d_ptr, unsigned char handle[ME_B200_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == ME_B200_IPC_HANDLE_BYTES, "IPC handle size");
  if (!ctx || !d_ptr || !handle) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  ME_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle, &h, sizeof h);
  return ME_OK;
}

int me_b200_ipc_open(me_b200_ctx *ctx, const unsigned char handle[ME_B200_IPC_HANDLE_BYTES], void **d_ptr) {
  if (!ctx || !handle || !d_ptr) return ME_ERR_INVALID_ARG;
  *d_ptr = nullptr;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  ME_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ME_OK;
}

int me_b200_ipc_close(me_b200_ctx *ctx, void *d_ptr) {
  if (!ctx || !d_ptr) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  ME_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
  return ME_OK;
}

int me_b200_search_device_band_peers(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                                     size_t pitch, size_t pair_stride, int npairs, int by_begin, int by_end,
                                     const me_b200_field *local, const me_b200_field *peers, int npeers,
                                     void *stream) {
  if (!ctx || !d_cur || !d_ref || !local || npairs < 1) return ME_ERR_INVALID_ARG;
  if (npeers < 0 || npeers >= ME_B200_MAX_PEERS || (npeers > 0 && !peers)) return ME_ERR_INVALID_ARG;
  if (pitch < (size_t)ctx->g.W) return ME_ERR_INVALID_ARG;
  if (npairs > 1 && pair_stride < pitch * (size_t)ctx->g.H) return ME_ERR_INVALID_ARG;
  if (by_begin < 0 || by_end > ctx->g.nby || by_begin > by_end) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  me::Frames f{d_cur, d_ref, pitch, pair_stride};
  me::Out o{local->mvx, local->mvy, local->ssd, local->score};
  me::Out po[ME_B200_MAX_PEERS];
  for (int i = 0; i < npeers; i++) po[i] = me::Out{peers[i].mvx, peers[i].mvy, peers[i].ssd, peers[i].score};
  // the tuned kernel stores into the peers itself; whatever block rows it did not cover (other
  // kernels, partial bottom rows) are stored by the small peer kernel afterwards
  if (ctx->plan) me::tiled_plan_set_peers(ctx->plan, po, npeers);
  rc = run_search(ctx, f, npairs, by_begin, by_end, o, (cudaStream_t)stream);
  int fb = by_begin, fe = by_begin;
  if (ctx->plan) {
    me::tiled_plan_fused_rows(ctx->plan, &fb, &fe);
    me::tiled_plan_set_peers(ctx->plan, nullptr, 0);
  }
  if (rc) return rc;
  if (npeers == 0) return ME_OK;
  if (fe <= fb) fb = fe = by_begin;
  cudaError_t e = me::launch_peer_scatter(ctx->g, npairs, by_begin, fb, o, po, npeers, (cudaStream_t)stream);
  if (e == cudaSuccess)
    e = me::launch_peer_scatter(ctx->g, npairs, fe, by_end, o, po, npeers, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_peer_scatter");
  if (fb > by_begin) ctx->launches++;
  if (by_end > fe) ctx->launches++;
  return ME_OK;
}

int me_b200_peer_barrier(me_b200_ctx *ctx, uint32_t *const *flags, int nranks, int my_rank, uint32_t epoch,
                         int timeout_ms, void *stream) {
  if (!ctx || !flags || nranks < 1 || nranks > ME_B200_MAX_PEERS || my_rank < 0 || my_rank >= nranks)
    return ME_ERR_INVALID_ARG;
  for (int i = 0; i < nranks; i++)
    if (!flags[i]) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  if (!ctx->d_peer_status) {
    ME_CUDA(ctx, cudaMalloc((void **)&ctx->d_peer_status, 256));
    ME_CUDA(ctx, cudaMemset(ctx->d_peer_status, 0, 256));
  }
  const unsigned long long ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 1000) * 1000000ull;
  cudaError_t e = me::launch_peer_barrier(flags, nranks, my_rank, epoch, ns, ctx->d_peer_status,
                                          (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_peer_barrier");
  ctx->launches++;
  return ME_OK;
}

int me_b200_peer_barrier_timed_out(me_b200_ctx *ctx, int *timed_out) {
  if (!ctx || !timed_out) return ME_ERR_INVALID_ARG;
  *timed_out = 0;
  if (!ctx->d_peer_status) return ME_OK;
  int rc = use_device(ctx);
  if (rc) return rc;
  ME_CUDA(ctx, cudaDeviceSynchronize());
  ME_CUDA(ctx, cudaMemcpy(timed_out, ctx->d_peer_status, sizeof(int), cudaMemcpyDeviceToHost));
  ME_CUDA(ctx, cudaMemset(ctx->d_peer_status, 0, sizeof(int)));
  return ME_OK;
}

int me_b200_search_device(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref, size_t pitch,
                          size_t pair_stride, int npairs, int32_t *d_mvx, int32_t *d_mvy,
                          uint32_t *d_ssd, float *d_score, void *stream) {
  if (!ctx) return ME_ERR_INVALID_ARG;
  return me_b200_search_device_band(ctx, d_cur, d_ref, pitch, pair_stride, npairs, 0, ctx->g.nby, d_mvx,
                                    d_mvy, d_ssd, d_score, stream);
}

int me_b200_postprocess_device(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref,
                               size_t pitch, const int32_t *d_mvx, const int32_t *d_mvy,
                               uint8_t *d_out5, unsigned long long *d_sq_err, uint32_t *d_max,
                               void *stream) {
  if (!ctx || !d_cur || !d_ref || !d_mvx || !d_mvy || !d_out5) return ME_ERR_INVALID_ARG;
  if (pitch < (size_t)ctx->g.W) return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaError_t e = me::launch_postprocess(ctx->g, d_cur, d_ref, pitch, d_mvx, d_mvy, d_out5, d_sq_err,
                                         d_max, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_postprocess");
  ctx->launches++;
  return ME_OK;
}

int me_b200_postprocess_device_batch(me_b200_ctx *ctx, const uint8_t *d_cur, const uint8_t *d_ref, size_t pitch,
                                     size_t pair_stride, int npairs, const int32_t *d_mvx, const int32_t *d_mvy,
                                     uint8_t *d_out5, size_t out_pair_stride, unsigned long long *d_sq_err,
                                     uint32_t *d_max, void *stream) {
  if (!ctx || !d_cur || !d_ref || !d_mvx || !d_mvy || !d_out5 || npairs < 1) return ME_ERR_INVALID_ARG;
  if (pitch < (size_t)ctx->g.W) return ME_ERR_INVALID_ARG;
  if (npairs > 1 && (pair_stride < pitch * (size_t)ctx->g.H || out_pair_stride < 5 * (size_t)ctx->g.W * ctx->g.H))
    return ME_ERR_INVALID_ARG;
  int rc = use_device(ctx);
  if (rc) return rc;
  cudaError_t e = me::launch_postprocess_batch(ctx->g, d_cur, d_ref, pitch, pair_stride, npairs, d_mvx, d_mvy, d_out5,
                                               out_pair_stride, d_sq_err, d_max, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_postprocess_batch");
  ctx->launches += (uint64_t)((npairs + 65534) / 65535);
  return ME_OK;
}

// ---- reference drop-in: int frames + predictionFrame -------------------------------

namespace {
struct cached_ctx {
  int device, W, H, B, R, cost, search;
  me_b200_ctx *ctx;
};
std::mutex g_cache_mu;
cached_ctx g_cache[8];
int g_cache_n = 0;
bool g_atexit = false;

int env_device() {
  const char *e = getenv("ME_B200_DEVICE");
  return e ? atoi(e) : 0;
}

extern "C" unsigned me_pack_int_to_u8(uint8_t *dst, const int *src, size_t n);  // host/me_pack.c

// Worker threads for the int -> u8 narrowing of the drop-in call (the reference keeps `int`
// pixels, utils.c:49-53).  A frame pair is cut into row chunks; workers narrow chunk after chunk
// into the pinned staging buffers and raise a per-chunk flag, the calling thread uploads every
// chunk as soon as its flag is up -- so narrowing and PCIe overlap and the call costs about
// max(narrow, upload) instead of their sum.  Workers spin for a short while after a job before
// they go back to sleep on the condition variable, which keeps back-to-back calls (a video) cheap.
constexpr int kMaxChunks = 32;

struct PackJob {
  const int *src[2];
  uint8_t *dst[2];
  size_t row_elems = 0;      // W
  int rows = 0, nchunks = 0; // chunks per frame; chunk c of frame f = flag index 2 * c + f
  int sub = 1;               // every chunk is narrowed as `sub` pieces, so that all workers share the FIRST
                             // chunk (the GPU starts when it has arrived) instead of one chunk each
  int row0[2][kMaxChunks + 1];  // row boundaries of the chunks, per frame (0 = reference, 1 = current)
};

class PackPool {
 public:
  explicit PackPool(int nthreads) {
    for (int i = 0; i < nthreads; i++) threads_.emplace_back([this] { worker(); });
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
  }
  // start narrowing `job`; chunk_ready(i) turns true chunk by chunk
  int workers() const { return (int)threads_.size(); }
  void start(const PackJob &job) {
    job_ = job;
    bad_.store(0, std::memory_order_relaxed);
    next_.store(0, std::memory_order_relaxed);
    for (int i = 0; i < 2 * job.nchunks; i++) done_[i].store(0, std::memory_order_relaxed);
    total_ = 2 * job.nchunks * job.sub;
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
  }
  // the caller helps (or does everything when there are no workers), then waits for flag i
  void wait_chunk(int i) {
    while (done_[i].load(std::memory_order_acquire) < job_.sub) {
      if (!run_one()) cpu_relax();
    }
  }
  unsigned bad_bits() const { return bad_.load(std::memory_order_acquire); }

 private:
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  bool run_one() {
    const int i = next_.fetch_add(1, std::memory_order_acq_rel);
    if (i >= total_) return false;
    // work item i = piece k of (chunk c, frame f), chunk-major so the first chunk is finished first
    const int k = i % job_.sub, cf = i / job_.sub, f = cf & 1, c = cf >> 1;
    const int r0 = job_.row0[f][c], nr = job_.row0[f][c + 1] - r0;
    const int p0 = r0 + (int)((long long)nr * k / job_.sub), p1 = r0 + (int)((long long)nr * (k + 1) / job_.sub);
    const size_t off = (size_t)p0 * job_.row_elems, n = (size_t)(p1 - p0) * job_.row_elems;
    const unsigned b = n ? me_pack_int_to_u8(job_.dst[f] + off, job_.src[f] + off, n) : 0u;
    if (b & ~0xffu) bad_.fetch_or(b, std::memory_order_relaxed);
    done_[cf].fetch_add(1, std::memory_order_release);
    return true;
  }
  void worker() {
    unsigned seen = 0;
    for (;;) {
      // spin briefly for the next job, then sleep
      unsigned g = gen_.load(std::memory_order_acquire);
      for (int spin = 0; g == seen && spin < 20000; spin++) {
        cpu_relax();
        g = gen_.load(std::memory_order_acquire);
      }
      if (g == seen) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
        g = gen_.load(std::memory_order_acquire);
      }
      seen = g;
      if (stop_) return;
      while (run_one()) {
      }
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<unsigned> gen_{0};
  bool stop_ = false;
  PackJob job_{};
  int total_ = 0;
  std::atomic<int> next_{1 << 30};
  std::atomic<unsigned> bad_{0};
  std::atomic<int> done_[2 * kMaxChunks];
};

PackPool *g_pack_pool = nullptr;  // created on the first drop-in call, under g_cache_mu

PackPool *pack_pool() {
  if (!g_pack_pool) {
    // measured on the 16-core B200 hosts (tools/dropin_sweep.py, 1080p): 6 workers 0.385 ms per call, 10 0.372, 14 0.366
    int n = (int)std::thread::hardware_concurrency() - 2;
    if (n > 10) n = 10;
    if (const char *e = getenv("ME_B200_PACK_THREADS")) n = atoi(e);
    if (n < 0) n = 0;
    if (n > 32) n = 32;
    g_pack_pool = new (std::nothrow) PackPool(n);
  }
  return g_pack_pool;
}
}  // namespace

void me_b200_release_cached(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (int i = 0; i < g_cache_n; i++) me_b200_destroy(g_cache[i].ctx);
  g_cache_n = 0;
  delete g_pack_pool;
  g_pack_pool = nullptr;
}

namespace {
int dropin_search(predictionFrame *pf, const int *refFrame, int extraSpan, int cost, int search, float *scores,
                  uint32_t *ssd) {
  if (!pf || !pf->frame || !pf->blks || !refFrame) return ME_ERR_INVALID_ARG;
  if (pf->width <= 0 || pf->height <= 0 || pf->blk_dim <= 0 || extraSpan < 0) return ME_ERR_INVALID_ARG;
  const int W = pf->width, H = pf->height, B = pf->blk_dim;
  const int nbx = (W + B - 1) / B, nby = (H + B - 1) / B;
  if (pf->num_blks != nbx * nby) return ME_ERR_UNSUPPORTED;
  // the grid must be the raster tiling createPredictionFrame builds (prediction_frame.c:14-23)
  for (int i = 0; i < pf->num_blks; i++) {
    const block &b = pf->blks[i];
    const int x0 = (i % nbx) * B, y0 = (i / nbx) * B;
    const int w = x0 + B < W ? B : W - x0, h = y0 + B < H ? B : H - y0;
    if (b.top_left_x != x0 || b.top_left_y != y0 || b.width != w || b.height != h)
      return ME_ERR_UNSUPPORTED;
  }
  const int device = env_device();
  std::lock_guard<std::mutex> lk(g_cache_mu);
  me_b200_ctx *ctx = nullptr;
  for (int i = 0; i < g_cache_n; i++)
    if (g_cache[i].device == device && g_cache[i].W == W && g_cache[i].H == H && g_cache[i].B == B &&
        g_cache[i].R == extraSpan && g_cache[i].cost == cost && g_cache[i].search == search)
      ctx = g_cache[i].ctx;
  if (!ctx) {
    int rc = me_b200_create_ex(&ctx, device, W, H, B, extraSpan, 1, ME_KERNEL_AUTO);
    if (rc) return rc;
    rc = me_b200_set_cost(ctx, cost);
    if (rc == ME_OK) rc = me_b200_set_search(ctx, search);
    if (rc) {
      me_b200_destroy(ctx);
      return rc;
    }
    if (g_cache_n == 8) {
      me_b200_destroy(g_cache[0].ctx);
      memmove(&g_cache[0], &g_cache[1], sizeof(cached_ctx) * 7);
      g_cache_n = 7;
    }
    g_cache[g_cache_n++] = cached_ctx{device, W, H, B, extraSpan, cost, search, ctx};
    if (!g_atexit) {
      g_atexit = true;
      atexit(me_b200_release_cached);
    }
  }
  int rc = use_device(ctx);
  if (rc) return rc;
  const size_t n = (size_t)W * H, nb = (size_t)ctx->nb;
  if (!ctx->h_cur) {
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_cur, n, cudaHostAllocDefault));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_ref, n, cudaHostAllocDefault));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_mvx, nb * 4, cudaHostAllocDefault));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_mvy, nb * 4, cudaHostAllocDefault));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_ssd, nb * 4, cudaHostAllocDefault));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_score, nb * 4, cudaHostAllocDefault));
  }
  // Narrow + upload + search as a pipeline over block-row BANDS of the frame: band c of the current
  // frame and the reference rows it needs (its own rows + R below; the R rows above arrived with the
  // band before) are narrowed by the workers (see PackPool), uploaded on the copy stream as soon as
  // their flag is up, and band c is searched on the compute stream behind an event -- while band c+1
  // is still being narrowed and uploaded.  The call costs about upload(first band) + search instead
  // of narrow + upload + search.  Small frames use one band (ME_B200_DROPIN_BANDS overrides).
  me_slot &sl = ctx->slots[0];
  cudaStream_t copy_stream = ctx->slots[1].stream;
  if (sl.busy || ctx->slots[1].busy) return ME_ERR_STATE;
  // ME_B200_TRACE=1: phase times of this call on stderr (host clock, microseconds since entry)
  static const bool trace = getenv("ME_B200_TRACE") != nullptr;
  const double t_entry = trace ? getTimeStamp() : 0.0;
  double t_band[kMaxChunks + 1] = {0}, t_launch[kMaxChunks + 1] = {0};
  PackPool *pool = pack_pool();
  if (!pool) return ME_ERR_NOMEM;
  // One launch for the whole frame while it is still arriving (tiled_plan_set_arrive) when the geometry
  // allows it -- then the bands are only the granularity of the arrival flag; otherwise one launch per band.
  bool arrive = n >= (1u << 20) && cost == ME_COST_MSE && search == ME_SEARCH_FULL && ctx->kernel == ME_KERNEL_TILED &&
                ctx->plan && me::tiled_arrive_supported(ctx->g);
  if (const char *e = getenv("ME_B200_DROPIN_ARRIVE")) arrive = arrive && e[0] != '0';
  int nbands = n >= (1u << 20) ? 4 : 1;
  if (const char *e = getenv("ME_B200_DROPIN_BANDS")) nbands = atoi(e);
  if (nbands < 1) nbands = 1;
  if (nbands > kMaxChunks) nbands = kMaxChunks;
  if (nbands > nby) nbands = nby;
  if (nbands < 2) arrive = false;
  if (!ctx->band_events[0]) {
    for (int i = 0; i < kMaxChunks; i++) ME_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_events[i], cudaEventDisableTiming));
  }
  if (arrive && !ctx->d_arrive) {
    ME_CUDA(ctx, cudaMalloc((void **)&ctx->d_arrive, 256));
    ME_CUDA(ctx, cudaMemset(ctx->d_arrive, 0, 256));
    ME_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_arrive, sizeof(unsigned int) * (kMaxChunks + 1), cudaHostAllocDefault));
  }
  // Arriving-frame launches send the REFERENCE frame first, whole (two pieces, so that the upload starts when half
  // of it is narrowed): once it is resident the energy-table pre-pass can run, and the search of the arriving
  // current frame is the table formulation (FORM 2 / 3: 203 instead of ~240 us for a 1080p pair) -- the reference
  // upload can cost less than that difference.  Otherwise the interleaved order: reference rows travel with the
  // band that needs them, FORM 1.
  // Default: 8x8 blocks only.  Same-box A/B (tools/jobs/r3h.sh): 4K 8x8 +-12 1.74 -> 1.55 ms per call, but 16x16 blocks
  // gain nothing (1080p +-32 0.370 vs 0.375, 4K +-32 1.27 vs 1.25 ms): FORM 1 is closer to FORM 3 there than the later
  // launch costs.  ME_B200_DROPIN_REF_FIRST=1 / 0 forces it on (any block size) / off.
  bool ref_first = arrive && B == 8;
  if (const char *e = getenv("ME_B200_DROPIN_REF_FIRST")) ref_first = arrive && e[0] != '0';
  const int nrefc = ref_first ? 2 : 0;            // chunks that carry only reference rows
  if (nrefc + nbands > kMaxChunks) nbands = kMaxChunks - nrefc;
  PackJob job;
  job.src[0] = refFrame; job.src[1] = pf->frame;
  job.dst[0] = ctx->h_ref; job.dst[1] = ctx->h_cur;
  job.row_elems = (size_t)W;
  job.rows = H;
  job.nchunks = nrefc + nbands;
  {
    // pieces of >= 64 KB (narrowed), one per worker + the calling thread at most
    int sub = pool->workers() + 1;
    const size_t band_bytes = n / (size_t)nbands;
    if ((size_t)sub > band_bytes / (64u << 10)) sub = (int)(band_bytes / (64u << 10));
    if (sub < 1) sub = 1;
    job.sub = sub;
  }
  int band_row[kMaxChunks + 1];   // block rows
  for (int c = 0; c <= nbands; c++) {
    band_row[c] = (int)((long long)nby * c / nbands);
    const int y = c == nbands ? H : (band_row[c] * B < H ? band_row[c] * B : H);
    job.row0[1][nrefc + c] = y;                                               // current frame: the band's own rows
    // reference: + R rows of halo below (reference first: nothing left for the band chunks)
    job.row0[0][nrefc + c] = ref_first ? H : (c == 0 ? 0 : (y + extraSpan < H ? y + extraSpan : H));
  }
  job.row0[0][nrefc + nbands] = H;
  for (int c = 0; c < nrefc; c++) {                                           // reference-only chunks
    job.row0[0][c] = (int)((long long)H * c / nrefc);
    job.row0[1][c] = 0;
  }
  unsigned int arrive_base = 0;
  if (arrive) {
    ctx->arrive_epoch = (ctx->arrive_epoch + 1) & 0x7fffu;
    if (ctx->arrive_epoch == 0) {
      // the 15-bit epoch wrapped: a flag left over from 32768 calls ago must not look like "rows resident"
      ctx->arrive_epoch = 1;
      ME_CUDA(ctx, cudaMemset(ctx->d_arrive, 0, sizeof(unsigned int)));
    }
    arrive_base = ctx->arrive_epoch << 16;
  }
  pool->start(job);
  cudaError_t ce = cudaSuccess;
  rc = ME_OK;
  me::Frames fr{sl.d_cur, sl.d_ref, ctx->pitch, ctx->frame_bytes};
  me::Out out{sl.d_mvx, sl.d_mvy, sl.d_ssd, sl.d_score};
  for (int c = 0; c < nbands && arrive; c++) ctx->h_arrive[c] = arrive_base + (unsigned int)job.row0[1][nrefc + c + 1];
  for (int cc = 0; cc < nrefc + nbands && ce == cudaSuccess && rc == ME_OK; cc++) {
    const int c = cc - nrefc;   // band index; negative: a reference-only chunk
    for (int f = 0; f < 2 && ce == cudaSuccess; f++) {
      pool->wait_chunk(2 * cc + f);
      const int r0 = job.row0[f][cc], nr = job.row0[f][cc + 1] - r0;
      if (nr <= 0) continue;
      uint8_t *d = (f == 0 ? sl.d_ref : sl.d_cur) + (size_t)r0 * ctx->pitch;
      const uint8_t *h = job.dst[f] + (size_t)r0 * W;
      if (ctx->pitch == (size_t)W)
        ce = cudaMemcpyAsync(d, h, (size_t)nr * W, cudaMemcpyHostToDevice, copy_stream);
      else
        ce = cudaMemcpy2DAsync(d, ctx->pitch, h, W, W, nr, cudaMemcpyHostToDevice, copy_stream);
    }
    if (trace) t_band[cc] = getTimeStamp();
    if (arrive) {
      // rows of band c (and the reference rows R below) are resident once this 4-byte copy has run
      if (c >= 0 && ce == cudaSuccess)
        ce = cudaMemcpyAsync(ctx->d_arrive, ctx->h_arrive + c, sizeof(unsigned int), cudaMemcpyHostToDevice, copy_stream);
      // the launch goes out behind the first band -- or, reference first, behind the last reference chunk (the
      // kernel then waits for the current frame's rows on the device)
      if (cc == (ref_first ? nrefc - 1 : 0) && ce == cudaSuccess) {
        ce = cudaEventRecord(ctx->band_events[0], copy_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sl.stream, ctx->band_events[0], 0);
        if (ce == cudaSuccess && !(pool->bad_bits() & ~0xffu)) {
          me::tiled_plan_set_arrive(ctx->plan, ctx->d_arrive, arrive_base, (int *)(ctx->d_arrive + 1), ref_first);
          rc = run_search(ctx, fr, 1, 0, nby, out, sl.stream);
        }
      }
      if (trace) t_launch[cc] = getTimeStamp();
      continue;
    }
    if (ce == cudaSuccess) ce = cudaEventRecord(ctx->band_events[c], copy_stream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(sl.stream, ctx->band_events[c], 0);
    if (ce == cudaSuccess && (pool->bad_bits() & ~0xffu)) break;   // a pixel outside 0..255: no point in searching
    if (ce == cudaSuccess) rc = run_search(ctx, fr, 1, band_row[c], band_row[c + 1], out, sl.stream);
    if (trace) t_launch[c] = getTimeStamp();
  }
  for (int i = 0; i < 2 * job.nchunks; i++) pool->wait_chunk(i);  // (after an error: let the workers finish)
  if (rc == ME_OK && ce != cudaSuccess) rc = fail_cuda(ctx, ce, "upload (host to device)");
  if (rc == ME_OK && (pool->bad_bits() & ~0xffu)) {
    snprintf(ctx->err, 256, "pixel value outside 0..255 (not representable in the 8-bit device layout)");
    rc = ME_ERR_UNSUPPORTED;
  }
  if (rc == ME_OK) {
    // only the arrays the caller asked for travel back
    const size_t ob = nb * 4;
    ce = cudaMemcpyAsync(ctx->h_mvx, sl.d_mvx, ob, cudaMemcpyDeviceToHost, sl.stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(ctx->h_mvy, sl.d_mvy, ob, cudaMemcpyDeviceToHost, sl.stream);
    if (ce == cudaSuccess && ssd) ce = cudaMemcpyAsync(ctx->h_ssd, sl.d_ssd, ob, cudaMemcpyDeviceToHost, sl.stream);
    if (ce == cudaSuccess && scores) ce = cudaMemcpyAsync(ctx->h_score, sl.d_score, ob, cudaMemcpyDeviceToHost, sl.stream);
    if (ce == cudaSuccess && arrive)   // the give-up status of the arriving-frame launch travels with the field
      ce = cudaMemcpyAsync(ctx->h_arrive + kMaxChunks, ctx->d_arrive + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost,
                           sl.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(sl.stream);
    if (ce != cudaSuccess) rc = fail_cuda(ctx, ce, "download (device to host)");
    if (rc == ME_OK && arrive && ctx->h_arrive[kMaxChunks]) {
      cudaMemset(ctx->d_arrive + 1, 0, sizeof(unsigned int));
      snprintf(ctx->err, 256, "the frame upload did not arrive while the search was waiting for it");
      rc = ME_ERR_CUDA;
    }
  }
  if (rc != ME_OK) {
    cudaStreamSynchronize(copy_stream);
    cudaStreamSynchronize(sl.stream);
    (void)cudaGetLastError();
  }
  if (trace) {
    const double t_done = getTimeStamp();
    fprintf(stderr, "[me_b200 trace] reference chunks %d, bands %d:", nrefc, nbands);
    for (int c = 0; c < nrefc + nbands; c++)
      fprintf(stderr, " [%d up %.0f launched %.0f]", c, (t_band[c] - t_entry) * 1e6, (t_launch[c] - t_entry) * 1e6);
    fprintf(stderr, " done %.0f us\n", (t_done - t_entry) * 1e6);
  }
  if (rc) {
    // the drop-in has no context handle to ask: its callers read me_b200_last_error(NULL)
    snprintf(g_err, 256, "%s", ctx->err);
    return rc;
  }
  for (int i = 0; i < pf->num_blks; i++) {
    block &b = pf->blks[i];
    b.motion_vectorX = ctx->h_mvx[i];  // populateBlkMotionVector, main.c:11-15
    b.motion_vectorY = ctx->h_mvy[i];
    b.is_best_match_found = 1;
  }
  if (scores) memcpy(scores, ctx->h_score, nb * 4);
  if (ssd) memcpy(ssd, ctx->h_ssd, nb * 4);
  return ME_OK;
}
}  // namespace

int me_b200_search_scores(predictionFrame *pf, const int *refFrame, int extraSpan, float *scores,
                          uint32_t *ssd) {
  return dropin_search(pf, refFrame, extraSpan, ME_COST_MSE, ME_SEARCH_FULL, scores, ssd);
}

int me_b200_search(predictionFrame *pf, const int *refFrame, int extraSpan) {
  return dropin_search(pf, refFrame, extraSpan, ME_COST_MSE, ME_SEARCH_FULL, nullptr, nullptr);
}

int me_b200_search_ssim_scores(predictionFrame *pf, const int *refFrame, int extraSpan, float *scores,
                               uint32_t *found) {
  return dropin_search(pf, refFrame, extraSpan, ME_COST_SSIM, ME_SEARCH_FULL, scores, found);
}

int me_b200_search_ssim(predictionFrame *pf, const int *refFrame, int extraSpan) {
  return dropin_search(pf, refFrame, extraSpan, ME_COST_SSIM, ME_SEARCH_FULL, nullptr, nullptr);
}

int me_b200_search_fast(predictionFrame *pf, const int *refFrame, int extraSpan, int search, float *scores,
                        uint32_t *ssd) {
  if (search != ME_SEARCH_THREE_STEP && search != ME_SEARCH_DIAMOND) return ME_ERR_INVALID_ARG;
  return dropin_search(pf, refFrame, extraSpan, ME_COST_MSE, search, scores, ssd);
}

}  // extern "C"
