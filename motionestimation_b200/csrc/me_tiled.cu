// me_tiled.cu -- tuned full-search kernel for sm_100a (the hot path).
//
// Reference being replaced: the per-block exhaustive scan of
// src/cpu/main.c:18-82 (computeMse / findBestMatchMse / findBestBlkMse) for all
// blocks of a batch of frame pairs (the dispatch loop main.c:144-158).
//
// Shape of the computation
//   An ITEM is a run of NS horizontally adjacent STRIPS of one block row of one
//   frame pair; a strip is NSUB adjacent blocks = 4*WORDS pixels wide.  All
//   candidates of an item live in one macro window of (NS*4*WORDS + 2R) x
//   (2R + BH) reference pixels.  A persistent CTA (one per SM, 16 warps, no
//   dedicated producer) walks its items through a 2..8 stage shared-memory ring:
//     * the macro window and the current-frame strip tile of an item arrive by TMA
//       (cp.async.bulk.tensor, 3-D u8 tensor maps, one mbarrier per stage).  TMA only
//       accepts 16-byte aligned inner coordinates (measured, tools/tma_probe.cu), so
//       the tile starts at the aligned column left of x0-R; frame borders cost
//       nothing: out-of-frame bytes are zero-filled and out-of-frame candidates are
//       never scored (main.c:73-76 clamps the window).
//     * warps pull 32-task chunks from a per-stage shared-memory counter.  A TASK is
//       (strip, horizontal offset dx, vertical part): the thread keeps the strip's
//       current-frame rows in registers (BH x WORDS words) and STREAMS down the
//       window column; each reference row (WORDS+1 aligned LDS.32, byte-aligned to
//       the candidate column by WORDS funnel shifts on the otherwise idle ALU pipe)
//       is scored against all BH current rows, feeding BH live candidates whose
//       accumulators rotate through a register file of BH slots (the period-BH
//       loop is fully unrolled so every index is static).  Four exact integer
//       formulations of the SSD are compiled (template FORM; 3 / 2 are the defaults):
//         FORM 1  SSD = sum cur^2 + sum ref^2 - 2 sum cur*ref: one IDP.4A.U8.U8
//                per 4 pixels for the cross term; sum ref^2 over the candidate's rows is a
//                sliding sum of per-row IDP.4A(ref,ref), sum cur^2 is per task.  Half the
//                instructions of FORM 0 (the unrolled loop fits the instruction cache),
//                FMA-pipe bound, and zero-padded current rows/columns make partial edge
//                blocks free (their reference pixels are masked out of sum ref^2).
//         FORM 2 (default)  as FORM 1, but sum ref^2 of every candidate position comes from a table that a
//                small pre-pass kernel (box_energy_kernel) builds per reference frame; the tile of
//                the table that belongs to an item rides in the stage next to the window (third
//                TMA load), so a finished candidate costs one LDS instead of 4 IDP.4A per row
//                step.  Used when the geometry has no partial-width blocks and the bigger stage
//                still fits twice; otherwise FORM 1.
//         FORM 3 (default for 16x16 blocks)  FORM 2 with the biased finish described below for 8x8
//                blocks: table = E + 2^24, ranking key t = table - 2*dot, per-thread key t << 7 | dy_rel,
//                per-block key t << 24 | dy << 16 | dx; sum cur^2 is added once per block at publish.
//         FORM 4 (opt-in, ME_B200_SSIM_FORM4=1)  the SSIM cost of me_ssim.cu on this kernel: the accumulators hold
//                sum r*c, the table holds {pixel sum, stddev} per reference rectangle, a finished candidate takes a
//                division-free bound and survivors are evaluated by all lanes together at the end of each period.
//                Bit-exact, but measured slower than me_ssim.cu's streaming kernel (profiles/ssim_form4_r02.txt).
//         FORM 0  |cur-ref| then square: VABSDIFF4.U8 (ALU pipe) + IDP.4A.U8.U8 (FMA pipe)
//                per 4 pixels; kept for A/B measurements (env ME_B200_FORM=0), full blocks only.
//     * a candidate that has seen its BH rows is folded into the thread's running
//       minimum of key32 = ssd << 8 | dy.  At the end of a chunk the lanes of one
//       block combine (MATCH.ANY + CREDUX.MIN) and one lane does a 64-bit shared
//       atomicMin of key32 << 32 | dx.  The unsigned minimum of (ssd, dy, dx) is
//       exactly the reference's first strict minimum in y-major/x-minor order
//       (main.c:53-62); ssd < 2^24 makes float(ssd)/float(w*h) the reference's
//       score bit for bit (main.c:19-27).
//     * the last warp to leave an item writes the item's motion vectors / SSD /
//       score (SoA), takes the next item from a launch-wide atomic counter (dynamic
//       scheduling over all CTAs) and re-arms the stage with its TMA loads.
//   8x8 blocks get one unrolled copy of the period per period kind (first / middle / last / single) instead of
//   warp-uniform branches inside one copy: a step holds only 32 IDP.4A there and the branches showed in the profile.
//   Vertical parts always hold m*BH + 1 candidates, so the ramp-up and ramp-down
//   of the rotating accumulators have a static shape and are skipped with
//   warp-uniform branches: no wasted pixel-compares, no validity tests in the loop.
//   Parts may overlap when the clamped window height is not of that form (a
//   duplicate candidate yields the same key, so the minimum is unchanged).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "me_device.cuh"
#include "me_ssim_math.cuh"
#include "me_tma.cuh"

namespace me {

namespace {

#ifndef ME_WARPS
#define ME_WARPS 16
#endif
#ifndef ME_PAIR_WARPS
#define ME_PAIR_WARPS 16   // pair kernel: 16 warps at 128 registers (a few spilled words outside the period loop) or 12 at 168
#endif
constexpr int kWarps = ME_WARPS;
constexpr int kThreads = kWarps * 32;
#ifndef ME_MAX_STAGES
#define ME_MAX_STAGES 8
#endif
constexpr int kMaxStages = ME_MAX_STAGES;
constexpr int kWinPitch = 256;  // window row pitch in shared memory = TMA box width (the maximum);
                                // a compile-time pitch turns every row offset into an LDS immediate
constexpr uint32_t kNoKey = 0xffffffffu;
constexpr int kMaxPeerOuts = 7;  // other GPUs of one NVSwitch domain
// FORM 2 with 8x8 blocks (64 pixels): the energy table holds E + kBias8, and a finished candidate is
// ranked by t = (E + kBias8) - 2*dot instead of the full SSD = sum cur^2 + E - 2*dot.  sum cur^2 is
// the same for every candidate of a block, so the order (and every tie) is unchanged, but the
// finish is LDS + IADD3 + PRMT + VIMNMX with nothing on the FMA-heavy pipe, and no task computes
// sum cur^2 at all; the publishing lane adds it back once per block.  2*dot <= 2*64*255^2 <
// kBias8 = 2^23 keeps t >= 0, and t <= 64*255^2 + 2^23 < 2^24 keeps it inside the 24-bit key field.
constexpr uint32_t kBias8 = 1u << 23;
// FORM 3 (16x16 blocks, 256 pixels): the same idea.  t = (E + kBias16) - 2*dot lies in (0, 2^25), so the
// per-thread key is t << 7 | dy relative to the first candidate row of the thread's vertical part (parts are
// limited to 113 candidates), and the shared 64-bit key is t << 24 | dy << 16 | dx.  No task computes
// sum cur^2 (64 IDP.4A per task on the saturated pipe, measured worth 1.7 %); the publishing lane adds it.
constexpr uint32_t kBias16 = 1u << 24;
constexpr int kMaxM16 = 7;   // FORM 3: part length m*16 + 1 <= 113 < 128
// FORM 4 (SSIM cost, 16x16 blocks): candidates that survive the division-free bound wait in a per-thread queue in
// shared memory (at most one push per step, drained at the end of every period of 16 steps)
constexpr int kQueueDepth = 16;

struct TiledParams {
  int W, H, B, R;
  int nbx, nby;           // blocks per row / rows of the whole frame
  int by_begin, by_count; // block rows handled by this launch
  int npairs;
  int strips_per_row;     // ceil(nbx / NSUB)
  int ns;                 // strips per item
  int items_per_row;      // ceil(strips_per_row / ns)
  int total_items;        // items_per_row * by_count * npairs
  int wb, wh;             // window box: bytes per row (pitch), rows
  int win_bytes;          // wb * wh rounded up to 128
  int cur_pitch;          // ns * 4 * WORDS bytes
  int stage_bytes;        // window + cur tile + best keys, 128-aligned
  int stages;             // ring depth, 2..kMaxStages
  int parts_target;       // wanted vertical parts per column
  int e;                  // bytes between the 16-aligned TMA origin and the window origin x0-R
  int skew;               // start-up stagger between the warps of one scheduler, in cycles
  int s_pitch, s_rows;    // FORM 2: energy tile, elements per row / rows (2R+1)
  int s_bytes;            // FORM 2: tile bytes rounded up to 128
  int e_s;                // FORM 2: elements between the 4-aligned TMA origin and x0-R
  int s_y0;               // FORM 2: frame row of the table's first row
  unsigned int inv_ndx;    // ceil(2^32 / (2R+1)): exact quotient by multiply-high for n < 2^16; 0 when R = 0
  unsigned int *next_item; // global work counter of this launch (zeroed on the stream before it)
  Out out;
  // band sharding over NVLink: every published block is also stored into the output arrays of the
  // peer GPUs (peer-mapped memory), so no collective follows the search
  int npeer;
  Out peer[kMaxPeerOuts];
  // "arriving frame" launches (the drop-in call): the frame pair is still being uploaded, top to bottom, while
  // the kernel runs.  *arrive_flag = arrive_base + number of current-frame rows that are resident (and the
  // reference rows R below them); an item waits for its rows before it starts its TMA loads.
  const unsigned int *arrive_flag;
  unsigned int arrive_base;
  int *arrive_status;   // set to 1 when an item gave up waiting
  // FORM 4 (SSIM cost): {pixel sum, stddev bits} of the current blocks of this launch, [pair][block row of the
  // launch][bx < nbx_full]; the pending-candidate queues start q_off bytes into the dynamic shared memory
  const int2 *blk_stats;
  int nbx_full;
  int q_off;
  // measurements only (env ME_B200_SSIM_STATS / ME_B200_SSIM_FAKE_THR): event counters of the drain loop, and a
  // made-up starting threshold (wrong results: what would perfect pruning be worth?)
  unsigned long long *dbg;
  unsigned int dbg_thr_bits;
};

// ---------------------------------------------------------------- item geometry
struct __align__(16) Item {
  int pair, by, strip0, ns;  // ns = strips actually present in this item
  int y0, h;                 // top pixel row of the block row, block height (BH or BH/2)
  int dy_lo, nc;             // first valid window-relative row offset, number of vertical candidates
  int m, nparts;             // part length = m*BH + 1
  int ntasks, nchunks, tpp;  // tpp = tasks per part
  unsigned inv_ns;           // ceil(2^32 / ns): row / ns by multiply-high (rows < 2^16)
  int pad_[2];               // 16 ints: four 16-byte shared-memory loads
};

template <int BH, int MAXM = 1 << 20>
__device__ __forceinline__ Item decode_item(const TiledParams &p, int it) {
  Item I;
  const int per_pair = p.items_per_row * p.by_count;
  I.pair = it / per_pair;
  int rem = it - I.pair * per_pair;
  const int row = rem / p.items_per_row;
  const int ir = rem - row * p.items_per_row;
  I.by = p.by_begin + row;
  I.strip0 = ir * p.ns;
  I.ns = min(p.ns, p.strips_per_row - I.strip0);
  I.y0 = I.by * p.B;
  I.h = min(BH, p.H - I.y0);
  // clamped window rows (main.c:74,76) as window-relative offsets dyr in [0, 2R], mvy = dyr - R
  I.dy_lo = max(0, p.R - I.y0);
  const int dy_hi = min(2 * p.R, p.H - I.h - I.y0 + p.R);
  I.nc = dy_hi - I.dy_lo + 1;
  const int want = (I.nc + p.parts_target - 1) / p.parts_target;
  I.m = min(min((want - 1 + BH - 1) / BH, (I.nc - 1) / BH), MAXM);
  const int L = I.m * BH + 1;
  I.nparts = (I.nc + L - 1) / L;
  I.tpp = I.ns * (2 * p.R + 1);
  I.inv_ns = (unsigned)((0x100000000ull + (unsigned)I.ns - 1) / (unsigned)I.ns);
  I.ntasks = I.tpp * I.nparts;
  I.nchunks = (I.ntasks + 31) >> 5;
  return I;
}

// period kind of the streaming loop as a compile-time tag or a run-time flag (see period_body)
template <bool V>
struct BoolTag {
  static constexpr bool value = V;
};
struct BoolRt {
  bool value;
};

// ME_TRACE_ITEMS (experiment builds only): per-item timeline in p.dbg[item * 8 + k], SM clock values:
//   0 re-arm starts  1 item index known (global atomic back)  2 TMA loads issued  3 first warp sees the data
//   4 last chunk handed out  5 last warp leaves (publish starts)  6 published   7 SM id
#ifdef ME_TRACE_ITEMS
#define ME_TRACE(item, k, how) do { if (p.dbg && (item) >= 0) how(p.dbg + (size_t)(item) * 8 + (k), (unsigned long long)clock64()); } while (0)
__device__ __forceinline__ void trace_set(unsigned long long *a, unsigned long long v) { *a = v; }
__device__ __forceinline__ void trace_min(unsigned long long *a, unsigned long long v) { atomicMin(a, v); }
__device__ __forceinline__ void trace_max(unsigned long long *a, unsigned long long v) { atomicMax(a, v); }
#else
#define ME_TRACE(item, k, how) do { } while (0)
#endif

// ---------------------------------------------------------------- the kernel
template <int WORDS, int BH, int NSUB, int FORM, bool PW, bool PEER, bool ARRIVE = false>
__global__ void __launch_bounds__(kThreads, 1)
tiled_search_kernel(const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_cur,
                    const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_sh,
                    const __grid_constant__ TiledParams p) {
  constexpr int SW = 4 * WORDS;       // strip width in pixels
  constexpr int BW = SW / NSUB;       // block width == p.B
  constexpr int WPB = WORDS / NSUB;   // words per block row
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];  // TMA bytes of the stage's item have landed
  // ring: next chunk of the stage's item.  queue: generation << 22 | chunks of the item << 11 | chunks handed out -- one atomicAdd
  // returns a consistent (generation, chunk count, chunk index) triple; a warp only waits for a stage's data after it
  // has claimed a chunk of that generation, so it can never wait for a phase that is two re-arms old.
  __shared__ uint32_t chunk_ctr[kMaxStages];
  __shared__ uint32_t left_ctr[kMaxStages];               // queue: finished chunks of the item; ring: warps that left it
  __shared__ int item_id[kMaxStages];                     // item held by the stage, -1 = no more work
  __shared__ Item item_s[kMaxStages];                     // its decoded geometry (five integer divisions: decoded
                                                          // once by the re-arming warp, not by all 16 warps)

  const int lane = threadIdx.x & 31;
  const int cur_off = p.win_bytes;                       // byte offsets inside a stage
  const int s_off = p.win_bytes + p.cur_pitch * BH;      // FORM 2: energy tile
  const int best_off = s_off + (FORM >= 2 ? p.s_bytes : 0);
  static_assert(FORM != 3 || (BH == 16 && NSUB == 1 && !PW && !PEER), "FORM 3 is the 16x16 table formulation");
  static_assert(FORM != 4 || (BH == 16 && NSUB == 1 && WORDS == 4 && !PW && !PEER && !ARRIVE), "FORM 4 is the 16x16 SSIM cost");
  constexpr int kSEntry = FORM == 4 ? 8 : 4;   // bytes per table entry (FORM 4: {pixel sum, stddev bits})
  // Stage protocol.  kQueue (8x8 blocks): the stages are a work queue -- a warp claims a chunk wherever one is left and
  // the warp that FINISHES an item's last chunk publishes it and re-arms the stage.  Otherwise (16x16): the lock-step
  // ring of round 1 -- every warp passes every item, the last one to leave re-arms.  Measured (late round 2, same
  // box): the queue is worth +5 % where items hold fewer chunks than the CTA has warps (4K 8x8 +-12 54.7 -> 57.4 %,
  // Foreman CIF 44.0 -> 46.5 %, 16x16 +-8 37.4 -> 39.6 %) and costs 1 to 2 % where chunks are long (16x16 +-32 81.2
  // -> 80.2 %, +-64 76.8 -> 75.0 %: two more shared-memory round trips per chunk), so each block size keeps the
  // protocol that is faster on the BASELINE configurations.
#ifndef ME_QUEUE_ALL
#define ME_QUEUE_ALL 0
#endif
  constexpr bool kQueue = BH == 8 || ME_QUEUE_ALL;
#ifndef ME_BALLOT_DX
#define ME_BALLOT_DX 1
#endif
  // chunk epilogue (me_tiled_chunk.inc): warp vote instead of a second masked reduction for the dx of the winner.
  // 8x8 kernels only: for FORM 3 (-DME_BALLOT_DX=2) it measured +1.5 % at +-8 / +-64 but -1.1 % on the headline
  // geometry (code placement), so 16x16 keeps the two reductions.
  constexpr bool kBallotDx = (ME_BALLOT_DX && BH == 8 && FORM != 4) || (ME_BALLOT_DX == 2 && FORM == 3);
  const int nblk_item = p.ns * NSUB;  // key slots per stage

  // (re)arm a stage: take the next item from the launch-wide counter (items are handed out in
  // row-major order, so the cheap clamped/half-height bottom rows come last and the tail is
  // short), reset the stage's keys and counters and start the TMA.  When the work is exhausted the
  // stage is marked dead and its barrier completed without a load.  Called by one whole warp.
  auto refill = [&](const int stage, const uint32_t gen) {
    uint8_t *sb = smem + (size_t)stage * p.stage_bytes;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sb + best_off);
    int it = 0;
#ifdef ME_TRACE_ITEMS
    const unsigned long long t_rearm = clock64();
#endif
    if (lane == 0) it = (int)atomicAdd(p.next_item, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
#ifdef ME_TRACE_ITEMS
    if (p.dbg && lane == 0 && it < p.total_items) {
      p.dbg[(size_t)it * 8 + 0] = t_rearm;
      p.dbg[(size_t)it * 8 + 1] = clock64();
      p.dbg[(size_t)it * 8 + 3] = ~0ull;
      p.dbg[(size_t)it * 8 + 4] = 0ull;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.dbg[(size_t)it * 8 + 7] = smid;
    }
#endif
    if (it >= p.total_items) {
      // queue: the hand-out word stays exhausted, nobody claims from, or waits for, a dead stage;
      // ring: the waiting warps are released by completing the phase without a load
      if (lane == 0) {
        *reinterpret_cast<volatile int *>(&item_id[stage]) = -1;
        if (!kQueue) mbar_arrive_expect_tx(&full_bar[stage], 0);
      }
      __syncwarp();
      return;
    }
    const Item I = decode_item<BH, (FORM == 3 ? kMaxM16 : (1 << 20))>(p, it);
    for (int b = lane; b < nblk_item; b += 32)   // FORM 4 keeps a maximum
      best[b] = FORM == 4 ? (unsigned long long)p.dbg_thr_bits << 32 : ~0ull;
    if (lane == 0) {
      if (!kQueue) chunk_ctr[stage] = 0;
      left_ctr[stage] = 0;
      item_id[stage] = it;
      item_s[stage] = I;
    }
    if constexpr (ARRIVE) {
      // the frame is still arriving: wait until the rows of this item are resident (the uploads and the
      // flag updates are ordered on one copy stream; items are handed out top to bottom)
      if (lane == 0) {
        const unsigned need = p.arrive_base + (unsigned)min(p.H, I.y0 + BH);
        for (int spins = 0;; spins++) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.arrive_flag) : "memory");
          if ((int)(v - need) >= 0) break;
          if (spins > (1 << 22)) {   // ~ 1 s: the upload failed or stalled; the host reports it
            *p.arrive_status = 1;
            break;
          }
          __nanosleep(200);
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      // order the generic-proxy accesses to this stage before the async-proxy (TMA) writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t bytes = (uint32_t)(p.wb * p.wh) + (uint32_t)(p.cur_pitch * BH);
      if (FORM >= 2) bytes += (uint32_t)(p.s_pitch * p.s_rows * kSEntry);
      mbar_arrive_expect_tx(&full_bar[stage], bytes);
      // 16-byte aligned origin: e bytes left of the window origin x0 - R
      tma_load_3d(sb, &map_ref, &full_bar[stage], I.strip0 * SW - p.R - p.e, I.y0 - p.R, I.pair);
      tma_load_3d(sb + cur_off, &map_cur, &full_bar[stage], I.strip0 * SW, I.y0, I.pair);
      if (FORM >= 2) {
        // energy tile: rows = window-relative dy 0..2R.  The half-height bottom row has its own
        // table whose row 0 is dy = 0 of that block row.
        if (FORM == 4 || I.h == BH)
          tma_load_3d(sb + s_off, &map_s, &full_bar[stage], I.strip0 * SW - p.R - p.e_s, I.y0 - p.R - p.s_y0, I.pair);
        else
          tma_load_3d(sb + s_off, &map_sh, &full_bar[stage], I.strip0 * SW - p.R - p.e_s, 0, I.pair);
      }
      ME_TRACE(it, 2, trace_set);
      if (kQueue) {
        // publish the item last: whoever claims a chunk from this word sees the geometry, the reset keys and counters
        __threadfence_block();
        atomicExch(&chunk_ctr[stage], ((gen & 0x3ffu) << 22) | ((uint32_t)I.nchunks << 11));
      }
    }
    __syncwarp();
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; s++) {
      mbar_init(&full_bar[s], 1);
      chunk_ctr[s] = 0;   // nothing to hand out yet
      item_id[s] = 0;     // ... but not dead either
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int k = 0; k < p.stages; k++) refill(k, 0u);

  // Optional start-up stagger of the warps that share a scheduler (env ME_B200_SKEW, cycles;
  // default 0 = off).  Measured: no effect on throughput -- the warps drift apart by themselves.
  if (p.skew > 0) {
    const long long until = clock64() + (long long)(threadIdx.x >> 7) * p.skew;
    while (clock64() < until) {
    }
  }

  // Two stage protocols (kQueue picks one per instantiation; the chunk work and the publish code are shared text,
  // me_tiled_chunk.inc / me_tiled_publish.inc):
  if constexpr (kQueue) {
  // The stages are a work queue, not a lock-step ring: a warp walks over them and claims a chunk wherever one is
  // left; the warp that FINISHES the last chunk of an item publishes it and re-arms the stage.  (Round 1 / early
  // round 2 re-armed a stage when all 16 warps had passed it -- with items of fewer chunks than warps every item
  // then waited for the slowest warp of the CTA: the per-item timeline, -DME_TRACE_ITEMS, showed stages waiting 2 to
  // 4 chunk times for their last visitor.)  A stage whose re-arm found no more items is dead; the walk ends when
  // all stages are dead.
  uint32_t dead = 0, idle_ns = 100u;
  bool progress = false;
  for (int stage = 0; dead != (1u << p.stages) - 1u;) {
    bool got = false;
    if (!(dead & (1u << stage))) {
      const uint32_t peek = *reinterpret_cast<volatile uint32_t *>(&chunk_ctr[stage]);
      if ((peek & 0x7ffu) < ((peek >> 11) & 0x7ffu)) {
        uint32_t word = 0;
        if (lane == 0) word = atomicAdd(&chunk_ctr[stage], 1u);
        word = __shfl_sync(0xffffffffu, word, 0);
        const int c = (int)(word & 0x7ffu), nch = (int)((word >> 11) & 0x7ffu);
        if (c < nch) {
          got = true;
          progress = true;
          // (no fence needed here: the geometry and the reset keys were written before the re-arming thread's
          // mbarrier arrive (release), and the try_wait below is the matching acquire)
    uint8_t *sb = smem + (size_t)stage * p.stage_bytes;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sb + best_off);
    mbar_wait(&full_bar[stage], (word >> 22) & 1u);
    const int it = *reinterpret_cast<volatile int *>(&item_id[stage]);
    const Item I = item_s[stage];
    if (lane == 0) ME_TRACE(it, 3, trace_min);
    if (lane == 0) ME_TRACE(it, 4, trace_max);

    const int L = I.m * BH + 1;
    const int ndx = 2 * p.R + 1;

    {
#include "me_tiled_chunk.inc"
    // ---- chunk finished; the warp that finishes the last one publishes the item and re-arms the stage
    __syncwarp();
    int prev = 0;
    if (lane == 0) {
      __threadfence_block();
      prev = (int)atomicAdd(&left_ctr[stage], 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == nch - 1) {
      __threadfence_block();
      if (lane == 0) ME_TRACE(it, 5, trace_set);
#include "me_tiled_publish.inc"
      __syncwarp();
      if (lane == 0) ME_TRACE(it, 6, trace_set);
      refill(stage, (word >> 22) + 1u);
    }
        }  // claimed a chunk
      } else if (*reinterpret_cast<volatile int *>(&item_id[stage]) < 0) {
        dead |= 1u << stage;
      }
    }
    if (!got) {
      // next stage; after a whole lap without a chunk, back off briefly (the others are computing or loading)
      // (exponential: a polling warp takes issue slots from the computing ones -- with a fixed 64 ns the long-chunk
      // geometries lost 1 to 2 %)
      if (++stage == p.stages) {
        stage = 0;
        if (!progress) {
          __nanosleep(idle_ns);
          idle_ns = idle_ns < 1600u ? idle_ns * 2u : idle_ns;
        } else {
          idle_ns = 100u;
        }
        progress = false;
      }
    }
  }
  } else {
  // Every warp walks the ring; a stage that reported "no more work" is never re-armed and is
  // skipped from then on; the walk ends when all stages are dead.
  uint32_t dead = 0;
  for (int k = 0; dead != (1u << p.stages) - 1u; k++) {
    const int stage = k % p.stages;
    if (dead & (1u << stage)) continue;
    uint8_t *sb = smem + (size_t)stage * p.stage_bytes;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sb + best_off);
    mbar_wait(&full_bar[stage], (k / p.stages) & 1);
    const int it = *reinterpret_cast<volatile int *>(&item_id[stage]);
    if (it < 0) {
      dead |= 1u << stage;
      continue;
    }
    const Item I = item_s[stage];
    if (lane == 0) ME_TRACE(it, 3, trace_min);

    const int L = I.m * BH + 1;
    const int ndx = 2 * p.R + 1;

    // (Tried for 8x8 blocks and dropped: fetching the NEXT chunk index while the current chunk is being worked on --
    // 1 to 4 % slower everywhere, the extra live value costs more than the hidden round trip.)
    for (;;) {
      int c = 0;
      if (lane == 0) c = (int)atomicAdd(&chunk_ctr[stage], 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= I.nchunks) break;
      if (lane == 0) ME_TRACE(it, 4, trace_max);

#include "me_tiled_chunk.inc"
    // ---- leave the item; the last warp out publishes it and re-arms the stage
    __syncwarp();
    int prev = 0;
    if (lane == 0) {
      __threadfence_block();
      prev = (int)atomicAdd(&left_ctr[stage], 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == kWarps - 1) {
      __threadfence_block();
      if (lane == 0) ME_TRACE(it, 5, trace_set);
#include "me_tiled_publish.inc"
      __syncwarp();
      if (lane == 0) ME_TRACE(it, 6, trace_set);
      refill(stage, 0u);
    }
  }
  }
}

// ---------------------------------------------------------------- 8x8 blocks: two block rows per item
// tiled_pair_kernel: the same search for 8x8 blocks, but an ITEM covers TWO vertically adjacent
// block rows and a TASK four blocks (two side by side x two on top of each other).  With 8x8
// blocks a task of the kernel above holds only 16 IDP.4A per candidate: its per-chunk set-up (task
// decode, current rows, the warp reduction and the shared atomics at the end) and the per-row
// fetch are amortised over too little work -- at +-12 the kernel is issue-bound at ~50 % of peak
// (profiles/ncu_tiled_8x8_pm12_r02_*).  Here the thread streams down ONE window column of
// 2R + 16 rows and scores every reference row against the 8 current rows of the upper block pair
// AND the 8 rows of the lower pair: the lower blocks' candidates simply lag the upper ones by one
// period (candidate dy of the lower block covers reference rows 8+dy .. 15+dy), so both use the
// same rotating-accumulator schedule, the same row fetch (5 LDS + 4 SHF per 64 IDP.4A instead of
// per 32) and the same energy-table row (two candidates that finish at the same step cover the
// same 8 reference rows: one LDS serves both).  A chunk carries twice the work for the same
// set-up.  64 current words + 32 accumulators need more registers than 512 threads leave, so the
// CTA has 12 warps (up to 168 registers per thread).
// Only block rows whose window is not clamped vertically are paired (the two rows of an item
// must have the same candidate rows); the few rows next to the top and bottom frame edge, an odd
// leftover row and every geometry FORM 2 does not cover run on the kernel above.
constexpr int kWarpsV = ME_PAIR_WARPS;

__device__ __forceinline__ Item decode_item_pair(const TiledParams &p, int it) {
  constexpr int BH = 8;
  Item I;
  const int per_pair = p.items_per_row * p.by_count;   // by_count: ITEM rows of this launch
  I.pair = it / per_pair;
  int rem = it - I.pair * per_pair;
  const int row = rem / p.items_per_row;
  const int ir = rem - row * p.items_per_row;
  I.by = p.by_begin + 2 * row;                         // the upper block row
  I.strip0 = ir * p.ns;
  I.ns = min(p.ns, p.strips_per_row - I.strip0);
  I.y0 = I.by * p.B;
  I.h = BH;
  I.dy_lo = 0;                                         // interior rows: the full 2R+1 candidate rows for both
  I.nc = 2 * p.R + 1;
  const int want = (I.nc + p.parts_target - 1) / p.parts_target;
  I.m = min((want - 1 + BH - 1) / BH, (I.nc - 1) / BH);
  const int L = I.m * BH + 1;
  I.nparts = (I.nc + L - 1) / L;
  I.tpp = I.ns * (2 * p.R + 1);
  I.inv_ns = (unsigned)((0x100000000ull + (unsigned)I.ns - 1) / (unsigned)I.ns);
  I.ntasks = I.tpp * I.nparts;
  I.nchunks = (I.ntasks + 31) >> 5;
  return I;
}

__global__ void __launch_bounds__(kWarpsV * 32, 1)
tiled_pair_kernel(const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_cur,
                  const __grid_constant__ CUtensorMap map_s, const __grid_constant__ TiledParams p) {
  constexpr int WORDS = 4, BH = 8, NSUB = 2, VS = 2;
  constexpr int SW = 4 * WORDS, BW = SW / NSUB, WPB = WORDS / NSUB, CR = BH * VS;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ uint32_t chunk_ctr[kMaxStages];
  __shared__ uint32_t left_ctr[kMaxStages];
  __shared__ int item_id[kMaxStages];
  __shared__ Item item_s[kMaxStages];

  const int lane = threadIdx.x & 31;
  const int cur_off = p.win_bytes;
  const int s_off = p.win_bytes + p.cur_pitch * CR;
  const int best_off = s_off + p.s_bytes;
  const int nblk_item = p.ns * NSUB * VS;   // key slot of (strip st, vertical v, sub-block b) = (st * VS + v) * NSUB + b

  auto refill = [&](const int stage) {
    uint8_t *sb = smem + (size_t)stage * p.stage_bytes;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sb + best_off);
    int it = 0;
    if (lane == 0) it = (int)atomicAdd(p.next_item, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= p.total_items) {
      if (lane == 0) {
        item_id[stage] = -1;
        mbar_arrive_expect_tx(&full_bar[stage], 0);
      }
      __syncwarp();
      return;
    }
    const Item I = decode_item_pair(p, it);
    for (int b = lane; b < nblk_item; b += 32) best[b] = ~0ull;
    if (lane == 0) {
      chunk_ctr[stage] = 0;
      left_ctr[stage] = 0;
      item_id[stage] = it;
      item_s[stage] = I;
    }
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t bytes = (uint32_t)(p.wb * p.wh) + (uint32_t)(p.cur_pitch * CR) + (uint32_t)(p.s_pitch * p.s_rows * 4);
      mbar_arrive_expect_tx(&full_bar[stage], bytes);
      tma_load_3d(sb, &map_ref, &full_bar[stage], I.strip0 * SW - p.R - p.e, I.y0 - p.R, I.pair);
      tma_load_3d(sb + cur_off, &map_cur, &full_bar[stage], I.strip0 * SW, I.y0, I.pair);
      // energy tile: row k = 8-row boxes whose top reference row is y0 - R + k, k = 0 .. 2R + 8
      tma_load_3d(sb + s_off, &map_s, &full_bar[stage], I.strip0 * SW - p.R - p.e_s, I.y0 - p.R - p.s_y0, I.pair);
    }
    __syncwarp();
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; s++) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32)
    for (int k = 0; k < p.stages; k++) refill(k);

  uint32_t dead = 0;
  for (int k = 0; dead != (1u << p.stages) - 1u; k++) {
    const int stage = k % p.stages;
    if (dead & (1u << stage)) continue;
    uint8_t *sb = smem + (size_t)stage * p.stage_bytes;
    unsigned long long *best = reinterpret_cast<unsigned long long *>(sb + best_off);
    mbar_wait(&full_bar[stage], (k / p.stages) & 1);
    const int it = *reinterpret_cast<volatile int *>(&item_id[stage]);
    if (it < 0) {
      dead |= 1u << stage;
      continue;
    }
    const Item I = item_s[stage];
    const int L = I.m * BH + 1;
    const int ndx = 2 * p.R + 1;

    for (;;) {
      int c = 0;
      if (lane == 0) c = (int)atomicAdd(&chunk_ctr[stage], 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= I.nchunks) break;

      int task = c * 32 + lane;
      const bool active = task < I.ntasks;
      task = min(task, I.ntasks - 1);
      const int row = p.inv_ndx ? (int)__umulhi((unsigned)task, p.inv_ndx) : task;  // = part * ns + strip
      const int dx = task - row * ndx;
      const int part = I.ns == 1 ? row : (int)__umulhi((unsigned)row, I.inv_ns);
      const int st = row - part * I.ns;
      const int u = p.e + st * SW + dx;
      const uint32_t shift = 8u * (uint32_t)(u & 3);
      const int c0 = I.nparts > 1 ? (int)(((long long)(I.nc - L) * part) / (I.nparts - 1)) : 0;

      // current rows of the four blocks: rows 0..7 = upper pair, 8..15 = lower pair
      uint32_t cur[CR][WORDS];
      {
        const uint4 *ct = reinterpret_cast<const uint4 *>(sb + cur_off);
        const int pitch4 = p.cur_pitch >> 4;
#pragma unroll
        for (int r = 0; r < CR; r++) {
          const uint4 v = ct[r * pitch4 + st];
          cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
        }
      }
      uint32_t acc[VS][NSUB][BH];
      uint32_t bestk[VS][NSUB];
#pragma unroll
      for (int v = 0; v < VS; v++)
#pragma unroll
        for (int b = 0; b < NSUB; b++) bestk[v][b] = kNoKey;

      const uint32_t *rowp = reinterpret_cast<const uint32_t *>(sb + (size_t)c0 * kWinPitch) + (u >> 2);
      constexpr int pitchw = kWinPitch >> 2;
      const int x_strip = (I.strip0 + st) * SW;
      int dy_fin = c0 - (BH - 1);   // dy of the UPPER candidate that finishes at step 0 of the current period

      uint32_t raw[WORDS + 1];
#pragma unroll
      for (int w = 0; w <= WORDS; w++) raw[w] = rowp[w];
      rowp += pitchw;

      const int m_uni = __reduce_max_sync(0xffffffffu, I.m);
      const int s_pitch = p.s_pitch;
      const uint32_t *spf = reinterpret_cast<const uint32_t *>(sb + s_off) + (c0 - (BH - 1)) * s_pitch + p.e_s + st * SW + dx;
      // periods 0 .. m+1: the upper blocks ramp up in period 0 and down in period m, the lower blocks one later
      for (int per = 0; per <= m_uni + 1; per++) {
        const bool u_lo = per < m_uni;                      // rows r < s_: candidates that started in this period exist
        const bool u_hi = per >= 1 && per <= m_uni;         // rows r > s_: candidates from the period before exist
        const bool u_dg = per <= m_uni;
        const bool l_lo = per >= 1 && per <= m_uni;
        const bool l_hi = per >= 2;
        const bool l_dg = per >= 1;
#pragma unroll
        for (int s_ = 0; s_ < BH; s_++) {
          uint32_t ref[WORDS];
#pragma unroll
          for (int w = 0; w < WORDS; w++) ref[w] = __funnelshift_r(raw[w], raw[w + 1], shift);
#pragma unroll
          for (int w = 0; w <= WORDS; w++) raw[w] = rowp[w];
          rowp += pitchw;

          // the energy of the 8 reference rows that end at this row: shared by the upper and the lower
          // candidate that finish now (only dereferenced when one of them exists)
          auto group = [&](const int v, const int r) {
            const int slot = (s_ - r + BH) % BH;
#pragma unroll
            for (int b = 0; b < NSUB; b++) {
              uint32_t a = (r == 0) ? 0u : acc[v][b][slot];
#pragma unroll
              for (int w = 0; w < WPB; w++) a = __dp4a(cur[v * BH + r][b * WPB + w], ref[b * WPB + w], a);
              acc[v][b][slot] = a;
              if (r == BH - 1) {
                const uint32_t ae = spf[b * BW];   // E + kBias8
                uint32_t t;
                asm("{ .reg .u32 t; sub.u32 t, %1, %2; sub.u32 %0, t, %2; }" : "=r"(t) : "r"(ae), "r"(a));
                const uint32_t key = __byte_perm(t, (uint32_t)(dy_fin + s_ - v * BH), 0x2104);
                bestk[v][b] = min(bestk[v][b], key);
              }
            }
          };
          if (u_lo) {
#pragma unroll
            for (int r = 0; r < s_; r++) group(0, r);
          }
          if (u_hi) {
#pragma unroll
            for (int r = s_ + 1; r < BH; r++) group(0, r);
          }
          if (l_lo) {
#pragma unroll
            for (int r = 0; r < s_; r++) group(1, r);
          }
          if (l_hi) {
#pragma unroll
            for (int r = s_ + 1; r < BH; r++) group(1, r);
          }
          if (u_dg) group(0, s_);
          if (l_dg) group(1, s_);
          spf += s_pitch;
        }
        dy_fin += BH;
      }

      const unsigned peers = __match_any_sync(0xffffffffu, active ? st : -1 - lane);
      const bool leader = (peers & (0u - peers)) == (1u << lane);
#pragma unroll
      for (int v = 0; v < VS; v++)
#pragma unroll
        for (int b = 0; b < NSUB; b++) {
          const int x0 = x_strip + b * BW;
          const bool ok = active && x0 < p.W && (x0 + dx - p.R >= 0) && (x0 + dx - p.R <= p.W - BW);
          const uint32_t key = ok ? bestk[v][b] : kNoKey;
          const uint32_t mkey = __reduce_min_sync(peers, key);
          const uint32_t mdx = __reduce_min_sync(peers, key == mkey ? (uint32_t)dx : 0xffffu);
          if (leader && mkey != kNoKey)
            atomicMin(&best[(st * VS + v) * NSUB + b], ((unsigned long long)mkey << 32) | mdx);
        }
    }
    __syncwarp();
    int prev = 0;
    if (lane == 0) {
      __threadfence_block();
      prev = (int)atomicAdd(&left_ctr[stage], 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == kWarpsV - 1) {
      __threadfence_block();
      for (int e = lane; e < I.ns * NSUB * VS; e += 32) {
        const int st = e / (NSUB * VS), rem = e - st * (NSUB * VS);
        const int v = rem / NSUB, b = rem - v * NSUB;
        const int bx = (I.strip0 + st) * NSUB + b;
        if (bx < p.nbx) {
          const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(&best[e]);
          const uint32_t k32 = (uint32_t)(key >> 32);
          // un-bias: ssd = t - kBias8 + sum cur^2 of this block, from the stage's current tile
          const uint32_t *ct = reinterpret_cast<const uint32_t *>(sb + cur_off) + (v * BH) * (p.cur_pitch >> 2) +
                               st * WORDS + b * WPB;
          uint32_t a2 = 0;
#pragma unroll
          for (int r = 0; r < BH; r++)
#pragma unroll
            for (int w = 0; w < WPB; w++) {
              const uint32_t x = ct[r * (p.cur_pitch >> 2) + w];
              a2 = __dp4a(x, x, a2);
            }
          const uint32_t ssd = (k32 >> 8) - kBias8 + a2;
          const size_t oi = (size_t)I.pair * p.nbx * p.nby + (size_t)(I.by + v) * p.nbx + bx;
          if (p.out.mvx) p.out.mvx[oi] = (int)(uint32_t)key - p.R;   // main.c:58
          if (p.out.mvy) p.out.mvy[oi] = (int)(k32 & 0xff) - p.R;    // main.c:59
          if (p.out.ssd) p.out.ssd[oi] = ssd;
          if (p.out.score) p.out.score[oi] = __fdiv_rn((float)ssd, (float)(BW * BH));  // main.c:27
        }
      }
      __syncwarp();
      refill(stage);
    }
  }
}

// ---------------------------------------------------------------- FORM 2 pre-pass
// box_energy_kernel: E(x, y) = sum of ref^2 over the bw x bh box whose top-left pixel is (x, y),
// for table rows y = y_lo .. y_lo + nrows - 1 and all x (boxes that leave the frame read zeros and
// are never used: such candidates do not exist).  One CTA = 128 x 40 table entries:
//   1. the pixels (+ halo) are staged in shared memory with 32-bit loads,
//   2. every aligned word gets its sum of 4 squares (one IDP.4A),
//   3. horizontal box sums: the 4-aligned position adds bw/4 word sums, the three positions after
//      it slide one pixel at a time (- leaving^2 + entering^2),
//   4. vertical box sums: one thread per column slides down the tile (+ entering row - leaving row)
//      and writes coalesced rows.
// Near HBM-bound (1 B read, 4 B written per entry); runs once per reference frame and launch.
constexpr int kEx = 128, kEy = 40, kEMax = 16;
constexpr int kEWords = (kEx + kEMax) / 4;  // aligned words per staged row

__global__ void __launch_bounds__(256)
box_energy_kernel(const uint8_t *__restrict__ ref, size_t pitch, size_t pair_stride, int W, int H, int bw, int bh,
                  int y_lo, int nrows, uint32_t *__restrict__ out, int out_pitch, size_t out_pair_stride,
                  uint32_t bias) {
  __shared__ uint32_t px[kEy + kEMax - 1][kEWords];   // pixels, 4 per word
  __shared__ uint32_t ws[kEy + kEMax - 1][kEWords];   // sum of squares of each aligned word
  __shared__ __align__(16) uint32_t hs[kEy + kEMax - 1][kEx];       // horizontal box sums
  const int x0 = blockIdx.x * kEx, r0 = blockIdx.y * kEy;  // r0: table row of this tile
  const uint8_t *src = ref + (size_t)blockIdx.z * pair_stride;
  const int rows_out = min(kEy, nrows - r0);
  const int rows_in = rows_out + bh - 1;
  const bool aligned = ((pitch & 3) == 0) && ((((uintptr_t)src) & 3) == 0);
  for (int i = threadIdx.x; i < rows_in * kEWords; i += 256) {
    const int r = i / kEWords, k = i - r * kEWords;
    const int y = y_lo + r0 + r, x = x0 + 4 * k;
    uint32_t w = 0;
    if (y >= 0 && y < H && x < W) {
      const uint8_t *q = src + (size_t)y * pitch + x;
      if (aligned && x + 4 <= W) {
        w = *reinterpret_cast<const uint32_t *>(q);
      } else {
        for (int b = 0; b < 4; b++)
          if (x + b < W) w |= (uint32_t)q[b] << (8 * b);
      }
    }
    px[r][k] = w;
    ws[r][k] = __dp4a(w, w, 0u);
  }
  __syncthreads();
  const int wpb = bw >> 2;  // aligned words per box (bw is 8 or 16)
  for (int i = threadIdx.x; i < rows_in * (kEx / 4); i += 256) {
    const int r = i / (kEx / 4), k = i - r * (kEx / 4);
    uint32_t a = 0;
    for (int j = 0; j < wpb; j++) a += ws[r][k + j];
    const uint32_t lo = px[r][k], hi = px[r][k + wpb];  // pixels leaving / entering as the box slides
    uint32_t o[4];
    o[0] = a;
#pragma unroll
    for (int b = 0; b < 3; b++) {
      const uint32_t l = (lo >> (8 * b)) & 0xffu, e = (hi >> (8 * b)) & 0xffu;
      a = a - l * l + e * e;
      o[b + 1] = a;
    }
    *reinterpret_cast<uint4 *>(&hs[r][4 * k]) = make_uint4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // vertical box sums: a thread owns four adjacent columns (16-byte loads and stores; a warp writes 512 contiguous
  // bytes per row) and one of eight row segments, all 256 threads busy
  static_assert(kEx / 4 == 32 && kEy % 8 == 0, "32 column groups x 8 row segments");
  constexpr int kSeg = kEy / 8;
  const int c4 = (threadIdx.x & 31) * 4, seg = threadIdx.x >> 5;
  const int rs = seg * kSeg, re = min(rs + kSeg, rows_out);
  if (rs < re && x0 + c4 < out_pitch) {   // (out_pitch is a multiple of 4: whole groups)
    uint32_t *dst = out + (size_t)blockIdx.z * out_pair_stride + (size_t)(r0 + rs) * out_pitch + x0 + c4;
    uint4 a = make_uint4(bias, bias, bias, bias);   // constant added to every entry (kBias8 / kBias16 / 0)
    for (int k = 0; k < bh; k++) {
      const uint4 v = *reinterpret_cast<const uint4 *>(&hs[rs + k][c4]);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<uint4 *>(dst) = a;
    for (int r = rs + 1; r < re; r++) {
      const uint4 in = *reinterpret_cast<const uint4 *>(&hs[r + bh - 1][c4]);
      const uint4 ou = *reinterpret_cast<const uint4 *>(&hs[r - 1][c4]);
      a.x += in.x - ou.x; a.y += in.y - ou.y; a.z += in.z - ou.z; a.w += in.w - ou.w;
      dst += out_pitch;
      *reinterpret_cast<uint4 *>(dst) = a;
    }
  }
}

// ---------------------------------------------------------------- host side
// block size -> kernel shape (WORDS, BH, NSUB) per formulation, or unsupported
bool shape_ok(int B) { return B == 16 || B == 8; }

}  // namespace

struct TiledPlan {
  int sms = 148;
  int max_smem = 0;
  int parts_target = 0;
  int ns_override = 0;
  unsigned long long kernels_launched = 0;  // every kernel this plan has launched (search + pre-pass)
  bool form_env_forced = false;  // ME_B200_FORM=2: use the table even for tiny launches (tests)
  // peer outputs of the NEXT launch (band sharding over NVLink), and the block rows that launch
  // actually wrote to the peers from inside the search kernel ([fused_begin, fused_end))
  Out peer[kMaxPeerOuts];
  int npeer = 0;
  int fused_begin = 0, fused_end = 0;
  bool fused_launch = false;  // the last search kernel launched was a peer-storing instantiation
  int form = 2;  // 2: energy table when possible, else 1 (default); 1: on-the-fly energies;
                 // 0: |a-b|^2 -- env ME_B200_FORM selects 0/1 for A/B measurements
  // stream-ordered scratch (energy tables, work counter) comes from the library's own pool
  // (scratch_pool): the device's default pool is shared with the host application and is left alone
  cudaMemPool_t pool = nullptr;
  // the last launch_tiled call enqueued work that writes the caller's outputs (false: the call failed
  // before that, so another kernel may serve the same request)
  bool outputs_enqueued = false;
  char err[160] = {0};
  // "arriving frame" mode of the NEXT launch (tiled_plan_set_arrive); consumed by that launch
  const unsigned int *arrive_flag = nullptr;
  unsigned int arrive_base = 0;
  int *arrive_status = nullptr;
  bool arrive_ref_resident = false;   // the whole reference frame is resident: the energy-table formulations may run
  // FORM 4 (SSIM cost; launch_tiled_ssim16): the statistics tables the caller built
  const int2 *ssim_table = nullptr;      // {pixel sum, stddev bits} per 16x16 rectangle of the reference frame
  int ssim_y_lo = 0, ssim_rows = 0;      // frame row of the table's first row, rows per pair
  size_t ssim_pair_stride = 0;           // entries per pair (row pitch = W entries)
  const int2 *ssim_blk = nullptr;        // the same for the current blocks of the launch
};

static int env_form() {
  const char *f = getenv("ME_B200_FORM");
  if (f && f[0] == '0') return 0;
  if (f && f[0] == '1') return 1;
  return 2;
}

bool tiled_supported(const Geom &g, size_t pitch, size_t pair_stride, const void *cur, const void *ref) {
  if (!shape_ok(g.B)) return false;
  if (env_form() == 0 && g.W % g.B != 0) return false;  // FORM 0 has no partial-width blocks
  if (g.R < 0 || 2 * g.R + g.B > 256) return false;  // key packs dy in 8 bits; TMA box <= 256 rows
  // one strip (16 px) plus the span plus the alignment slack must fit the 256-byte window row
  if ((16 - g.R % 16) % 16 + 16 + 2 * g.R + 4 > kWinPitch) return false;
  if (g.W < g.B || g.H < g.B) return false;
  if ((pitch & 15) || (pair_stride & 15)) return false;  // TMA: 16-byte aligned base and strides
  if (((uintptr_t)cur & 15) || ((uintptr_t)ref & 15)) return false;
  if (!get_encode()) return false;
  return true;
}

cudaError_t tiled_plan_create(TiledPlan **plan, const Geom &g, int /*max_pairs*/) {
  *plan = nullptr;
  if (!shape_ok(g.B)) return cudaErrorNotSupported;
  TiledPlan *pl = new TiledPlan();
  pl->form = env_form();
  {
    const char *f = getenv("ME_B200_FORM");
    pl->form_env_forced = f && f[0] == '2';
  }

  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) {
    delete pl;
    return e;
  }
  e = scratch_pool(&pl->pool);
  if (e != cudaSuccess) {
    delete pl;
    return e;
  }
  const char *pt = getenv("ME_B200_PARTS");
  const char *ns = getenv("ME_B200_NS");
  pl->parts_target = pt ? atoi(pt) : 0;
  pl->ns_override = ns ? atoi(ns) : 0;
  *plan = pl;
  return cudaSuccess;
}

void tiled_plan_destroy(TiledPlan *plan) {
  delete plan;
}

int tiled_plan_set_peers(TiledPlan *plan, const Out *peers, int npeers) {
  if (!plan || npeers < 0 || npeers > kMaxPeerOuts) return -1;
  for (int i = 0; i < npeers; i++) plan->peer[i] = peers[i];
  plan->npeer = npeers;
  plan->fused_begin = plan->fused_end = 0;
  plan->fused_launch = false;
  return 0;
}

void tiled_plan_set_arrive(TiledPlan *plan, const unsigned int *flag, unsigned int base, int *status,
                           bool ref_resident) {
  if (!plan) return;
  plan->arrive_ref_resident = ref_resident;
  plan->arrive_flag = flag;
  plan->arrive_base = base;
  plan->arrive_status = status;
}

bool tiled_arrive_supported(const Geom &g) {
  // every block row must run on the tuned kernel (a generic tail launch would need the whole frame), and
  // the on-the-fly energy formulation is the one without a pre-pass over the whole reference frame
  if (!shape_ok(g.B) || env_form() == 0 || g.H >= 65536) return false;
  const int hrem = g.H % g.B;
  return hrem == 0 || hrem == g.B / 2;
}

void tiled_plan_fused_rows(const TiledPlan *plan, int *begin, int *end) {
  *begin = plan ? plan->fused_begin : 0;
  *end = plan ? plan->fused_end : 0;
}
unsigned long long tiled_plan_launches(const TiledPlan *plan) { return plan ? plan->kernels_launched : 0; }
bool tiled_plan_outputs_enqueued(const TiledPlan *plan) { return plan && plan->outputs_enqueued; }

namespace {

// VS = 2: items of two block rows (tiled_pair_kernel; 8x8 blocks, FORM 2, by_count even, rows whose
// window is not clamped vertically)
template <int WORDS, int BH, int NSUB, int FORM, bool PW, int VS = 1>
cudaError_t launch_shape_pw(TiledPlan *plan, const Geom &g, const Frames &f, int npairs, const Out &o,
                         int by_begin, int by_count, cudaStream_t s, const char **err) {
  static_assert(VS == 1 || (VS == 2 && FORM == 2 && BH == 8 && NSUB == 2 && WORDS == 4 && !PW), "pair kernel shape");
  constexpr int SW = 4 * WORDS;
  constexpr int kW = VS == 2 ? kWarpsV : kWarps;   // warps per CTA
  TiledParams p;
  memset(&p, 0, sizeof p);
  p.W = g.W; p.H = g.H; p.B = g.B; p.R = g.R;
  p.nbx = g.nbx; p.nby = g.nby;
  p.by_begin = by_begin; p.by_count = by_count / VS;   // item rows
  p.npairs = npairs;
  p.strips_per_row = FORM == 4 ? g.W / SW : (g.nbx + NSUB - 1) / NSUB;   // FORM 4: blocks of full width only
  p.wh = 2 * g.R + BH * VS;
  const int static_smem = 2048;  // static shared memory (barriers, counters) + alignment slack
  const int ebytes = (16 - (g.R % 16)) % 16;  // SW is a multiple of 16, so every item has the same phase
  // Choose strips per item (ns) and vertical parts per column with a small cost model:
  // a CTA's time ~ (items it owns) x (chunks per item) x (instructions per task) / warps.
  // ns is bounded by the TMA box (<= 256 B per row) and by two stages of shared memory.
  const long long rows_total = (long long)(by_count / VS) * npairs;
  const int nc = 2 * g.R + 1;
  double best_cost = 1e300;
  int best_ns = 0, best_parts = 1;
  constexpr int BWc = SW / NSUB;
  // table tile: the TMA origin must be 16-byte aligned (4 u32 entries; FORM 4: 2 entries of 8 bytes);
  // SW is a multiple of 4 elements, so every item has the same phase
  constexpr int kSEntry = FORM == 4 ? 8 : 4;
  constexpr int kSAlign = 16 / kSEntry;
  const int e_s = (kSAlign - (g.R % kSAlign)) % kSAlign;
  auto s_pitch_of = [&](int ns) { return (e_s + ns * SW + 2 * g.R - BWc + 1 + kSAlign - 1) & ~(kSAlign - 1); };
  const int queue_bytes = FORM == 4 ? kThreads * kQueueDepth * 4 : 0;
  auto stage_size = [&](int ns) {
    const int wb = kWinPitch;
    const int win = ((wb * p.wh) + 127) & ~127;
    const int stile = FORM >= 2 ? ((s_pitch_of(ns) * (2 * g.R + 1 + BH * (VS - 1)) * kSEntry + 127) & ~127) : 0;
    return (win + ns * SW * BH * VS + stile + ns * NSUB * VS * 8 + 127) & ~127;
  };
  for (int ns = 1; ns <= p.strips_per_row && ns <= 16; ns++) {
    const int wb = (ebytes + ns * SW + 2 * g.R + 4 + 15) & ~15;
    if (wb > 256 || ns * SW > 256) break;
    if (FORM >= 2 && s_pitch_of(ns) > 256) break;
    if (2 * stage_size(ns) + static_smem + queue_bytes > plan->max_smem) break;
    if (plan->ns_override > 0 && ns != plan->ns_override) continue;
    const long long items = rows_total * ((p.strips_per_row + ns - 1) / ns);
    const long long per_cta = (items + plan->sms - 1) / plan->sms;
    for (int parts = 1; parts <= 8; parts++) {
      if (plan->parts_target > 0 && parts != plan->parts_target) continue;
      const int want = (nc + parts - 1) / parts;
      int m = (want - 1 + BH - 1) / BH;
      if (m > (nc - 1) / BH) m = (nc - 1) / BH;
      if (FORM == 3 && m > kMaxM16) m = kMaxM16;   // as decode_item does
      const int L = m * BH + 1;
      const int nparts = (nc + L - 1) / L;
      const long long chunks = ((long long)ns * nc * nparts + 31) / 32;
      if (chunks > 2000) continue;   // the stage queue's hand-out word holds 11 bits of chunk count (+ 16 of overshoot)
      // per task: L candidates x BH rows x WORDS cross-term ops (x2 for FORM 0), plus per streamed
      // row the loads, shifts and (FORM 1) the row-energy ops
      const double per_row = FORM == 1 ? (2.0 * WORDS + 8.0) : (2.0 * WORDS + 6.0);
      const double task = (double)L * BH * VS * WORDS * (FORM >= 1 ? 1.0 : 2.0) + (double)(L + BH * VS - 1) * per_row + 250.0;
      // warps flow from one item into the next, so chunks only quantise over the CTA's whole run
      // + per item: every warp's barrier wait / item decode / counter round trips, the re-arm and the publish
      const double cost = (double)((per_cta * chunks + kW - 1) / kW) * task + 300.0 * per_cta;
      if (cost < best_cost * 0.999) { best_cost = cost; best_ns = ns; best_parts = parts; }
    }
  }
  if (best_ns == 0) { *err = "tiled: window does not fit shared memory"; return cudaErrorInvalidConfiguration; }
  const int ns = best_ns;
  p.ns = ns;
  p.parts_target = best_parts;
  p.items_per_row = (p.strips_per_row + ns - 1) / ns;
  p.total_items = (int)(p.items_per_row * rows_total);
  p.wb = kWinPitch;
  p.e = ebytes;
  p.win_bytes = ((p.wb * p.wh) + 127) & ~127;
  p.cur_pitch = ns * SW;
  p.s_pitch = s_pitch_of(ns);
  p.s_rows = 2 * g.R + 1 + BH * (VS - 1);
  p.s_bytes = (p.s_pitch * p.s_rows * kSEntry + 127) & ~127;
  p.e_s = e_s;
  p.stage_bytes = stage_size(ns);
  p.stages = (plan->max_smem - static_smem - queue_bytes) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.out = o;
  // peer stores from inside the kernel exist for the two default formulations on full-width
  // frames; other launches leave the peers to the caller's store kernel
  constexpr bool kPeerVariant = FORM >= 1 && FORM < 3 && !PW && VS == 1;
  const bool peer = kPeerVariant && plan->npeer > 0;
  p.npeer = peer ? plan->npeer : 0;
  for (int q = 0; q < p.npeer; q++) p.peer[q] = plan->peer[q];
  p.inv_ndx = g.R == 0 ? 0u
                       : (unsigned int)((0x100000000ull + (unsigned)(2 * g.R + 1) - 1) / (unsigned)(2 * g.R + 1));
  {
    const char *sk = getenv("ME_B200_SKEW");
    p.skew = sk ? atoi(sk) : 0;
  }

  EncodeTiledFn enc = get_encode();
  if (!enc) { *err = "cuTensorMapEncodeTiled unavailable"; return cudaErrorNotSupported; }
  CUtensorMap map_ref, map_cur;
  const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)npairs};
  const cuuint64_t strides[2] = {(cuuint64_t)f.pitch, (cuuint64_t)(npairs > 1 ? f.pair_stride : f.pitch * g.H)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t box_ref[3] = {(cuuint32_t)p.wb, (cuuint32_t)p.wh, 1};
  const cuuint32_t box_cur[3] = {(cuuint32_t)p.cur_pitch, (cuuint32_t)(BH * VS), 1};
  CUresult r1 = enc(&map_ref, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)f.ref, dims, strides, box_ref, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&map_cur, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)f.cur, dims, strides, box_cur, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
    snprintf(plan->err, sizeof plan->err, "cuTensorMapEncodeTiled failed (%d, %d) wb=%d wh=%d", (int)r1, (int)r2,
             p.wb, p.wh);
    *err = plan->err;
    return cudaErrorInvalidValue;
  }
  // FORM 2: build the energy tables for the rows this launch needs (stream-ordered scratch)
  CUtensorMap map_s = map_ref, map_sh = map_ref;  // placeholders for the other formulations
  uint32_t *d_s = nullptr;
  if constexpr (FORM == 4) {
    // SSIM cost: the caller (me_ssim.cu) built the statistics tables; entries are 8 bytes, row pitch = W entries
    p.s_y0 = plan->ssim_y_lo;
    p.blk_stats = plan->ssim_blk;
    p.nbx_full = g.W / SW;
    const cuuint64_t sd[3] = {(cuuint64_t)g.W, (cuuint64_t)plan->ssim_rows, (cuuint64_t)npairs};
    const cuuint64_t sstr[2] = {(cuuint64_t)g.W * 8, (cuuint64_t)plan->ssim_pair_stride * 8};
    const cuuint32_t sbox[3] = {(cuuint32_t)p.s_pitch, (cuuint32_t)p.s_rows, 1};
    CUresult r3 = enc(&map_s, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)plan->ssim_table, sd, sstr, sbox, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r3 != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled(SSIM statistics table) failed"; return cudaErrorInvalidValue; }
    map_sh = map_s;
    if (const char *ft = getenv("ME_B200_SSIM_FAKE_THR")) {
      const float v = (float)atof(ft);
      memcpy(&p.dbg_thr_bits, &v, 4);
    }
    if (getenv("ME_B200_SSIM_STATS")) {
      if (cudaMalloc((void **)&p.dbg, 64) == cudaSuccess) cudaMemset(p.dbg, 0, 64);
      else p.dbg = nullptr;
    }
  }
  if (FORM == 2 || FORM == 3) {
    constexpr int BW = SW / NSUB;
    const int tp = (g.W + 3) & ~3;  // table pitch in elements
    const int full_rows = g.H / g.B;
    const int last_full = (by_begin + by_count < full_rows ? by_begin + by_count : full_rows) - 1;
    int y_lo = by_begin * g.B - g.R, y_hi = last_full * g.B + g.R;
    if (y_lo < 0) y_lo = 0;
    if (y_hi > g.H - BH) y_hi = g.H - BH;
    const int nfull = (by_begin < full_rows && y_hi >= y_lo) ? y_hi - y_lo + 1 : 0;
    const bool has_half = by_begin + by_count > full_rows;  // the bottom row of height BH/2
    const int nhalf = has_half ? g.R + 1 : 0;
    const size_t per_pair = (size_t)tp * (size_t)(nfull + nhalf);
    cudaError_t e = cudaMallocFromPoolAsync((void **)&d_s, per_pair * 4 * (size_t)npairs + 256, plan->pool, s);
    if (e != cudaSuccess) { *err = "cudaMallocAsync(energy table)"; return e; }
    const size_t ref_pair_stride = npairs > 1 ? f.pair_stride : f.pitch * g.H;
    if (nfull > 0) {
      dim3 eg((tp + kEx - 1) / kEx, (nfull + kEy - 1) / kEy, npairs);
      box_energy_kernel<<<eg, 256, 0, s>>>(f.ref, f.pitch, ref_pair_stride, g.W, g.H, BW, BH, y_lo, nfull, d_s, tp,
                                           per_pair, BH == 8 ? kBias8 : (FORM == 3 ? kBias16 : 0u));
      plan->kernels_launched++;
    }
    if (nhalf > 0) {
      dim3 eg((tp + kEx - 1) / kEx, (nhalf + kEy - 1) / kEy, npairs);
      box_energy_kernel<<<eg, 256, 0, s>>>(f.ref, f.pitch, ref_pair_stride, g.W, g.H, BW, BH / 2,
                                           g.H - BH / 2 - g.R, nhalf, d_s + (size_t)tp * nfull, tp, per_pair,
                                           BH == 8 ? kBias8 : (FORM == 3 ? kBias16 : 0u));
      plan->kernels_launched++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = "box_energy_kernel launch"; cudaFreeAsync(d_s, s); return e; }
    p.s_y0 = y_lo;
    const cuuint64_t sstr[2] = {(cuuint64_t)tp * 4, (cuuint64_t)per_pair * 4};
    const cuuint32_t sbox[3] = {(cuuint32_t)p.s_pitch, (cuuint32_t)p.s_rows, 1};
    CUresult r3 = CUDA_SUCCESS, r4 = CUDA_SUCCESS;
    if (nfull > 0) {
      const cuuint64_t sd[3] = {(cuuint64_t)tp, (cuuint64_t)nfull, (cuuint64_t)npairs};
      r3 = enc(&map_s, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)d_s, sd, sstr, sbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (nhalf > 0) {
      const cuuint64_t sd[3] = {(cuuint64_t)tp, (cuuint64_t)nhalf, (cuuint64_t)npairs};
      r4 = enc(&map_sh, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)(d_s + (size_t)tp * nfull), sd, sstr, sbox, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r3 != CUDA_SUCCESS || r4 != CUDA_SUCCESS) {
      *err = "cuTensorMapEncodeTiled(energy table) failed";
      cudaFreeAsync(d_s, s);
      return cudaErrorInvalidValue;
    }
    if (nfull == 0) map_s = map_sh;
    if (nhalf == 0) map_sh = map_s;
  }
  // launch-wide work counter (stream-ordered scratch, zeroed on the stream)
  unsigned int *d_ctr = nullptr;
  {
    cudaError_t ce = cudaMallocFromPoolAsync((void **)&d_ctr, 256, plan->pool, s);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_ctr, 0, 256, s);
    if (ce != cudaSuccess) {
      *err = "cudaMallocAsync(work counter)";
      if (d_s) cudaFreeAsync(d_s, s);
      if (d_ctr) cudaFreeAsync(d_ctr, s);
      return ce;
    }
    p.next_item = d_ctr;
  }
#ifdef ME_TRACE_ITEMS
  if (FORM != 4 && getenv("ME_B200_TRACE_ITEMS")) {
    if (cudaMalloc((void **)&p.dbg, (size_t)p.total_items * 64) == cudaSuccess) cudaMemset(p.dbg, 0, (size_t)p.total_items * 64);
    else p.dbg = nullptr;
  }
#endif
  p.q_off = p.stages * p.stage_bytes;
  const int smem = p.stages * p.stage_bytes + queue_bytes;
  if (getenv("ME_B200_VERBOSE"))
    fprintf(stderr, "[me_b200] tiled<%d,%d,%d,form %d> ns=%d parts=%d items=%d stages=%d stage=%d B smem=%d B s_pitch=%d\n",
            WORDS, BH, NSUB, FORM, p.ns, p.parts_target, p.total_items, p.stages, p.stage_bytes, smem, p.s_pitch);
  const int grid = p.total_items < plan->sms ? p.total_items : plan->sms;
  cudaError_t e;
  if constexpr (VS == 2) {
    e = cudaFuncSetAttribute(tiled_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) tiled_pair_kernel<<<grid, kWarpsV * 32, smem, s>>>(map_ref, map_cur, map_s, p);
  } else {
    auto kern = tiled_search_kernel<WORDS, BH, NSUB, FORM, PW, false>;
    if constexpr (kPeerVariant) {
      if (peer) kern = tiled_search_kernel<WORDS, BH, NSUB, FORM, PW, true>;
    }
    // arriving-frame instantiations: FORM 1 (any geometry), and the table formulations for full-width frames whose
    // reference frame is already resident (the pre-pass above ran on it)
    if constexpr (FORM == 1 || ((FORM == 3 || (FORM == 2 && BH == 8)) && !PW)) {
      if (plan->arrive_flag && !peer) {
        p.arrive_flag = plan->arrive_flag;
        p.arrive_base = plan->arrive_base;
        p.arrive_status = plan->arrive_status;
        kern = tiled_search_kernel<WORDS, BH, NSUB, FORM, PW, false, true>;
      }
    }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) kern<<<grid, kThreads, smem, s>>>(map_ref, map_cur, map_s, map_sh, p);
  }
  if (e != cudaSuccess) {
    *err = "cudaFuncSetAttribute(tiled)";
    if (d_s) cudaFreeAsync(d_s, s);
    cudaFreeAsync(d_ctr, s);
    return e;
  }
  plan->kernels_launched++;
  e = cudaGetLastError();
#ifdef ME_TRACE_ITEMS
  if (FORM != 4 && p.dbg) {
    cudaStreamSynchronize(s);
    const size_t n = (size_t)p.total_items;
    unsigned long long *h = (unsigned long long *)malloc(n * 64);
    cudaMemcpy(h, p.dbg, n * 64, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    double d[6] = {0, 0, 0, 0, 0, 0};
    size_t cnt = 0;
    unsigned long long sm_first[256], sm_last[256];
    for (int i = 0; i < 256; i++) { sm_first[i] = ~0ull; sm_last[i] = 0; }
    for (size_t i = 0; i < n; i++) {
      const unsigned long long *t = h + i * 8;
      if (!t[6] || t[3] == ~0ull) continue;
      d[0] += (double)(t[1] - t[0]); d[1] += (double)(t[2] - t[1]); d[2] += (double)((long long)(t[3] - t[2]));
      d[3] += (double)((long long)(t[4] - t[3])); d[4] += (double)((long long)(t[5] - t[4])); d[5] += (double)(t[6] - t[5]);
      const int sm = (int)(t[7] & 255);
      if (t[0] < sm_first[sm]) sm_first[sm] = t[0];
      if (t[6] > sm_last[sm]) sm_last[sm] = t[6];
      cnt++;
    }
    double span = 0; int nsm = 0;
    for (int i = 0; i < 256; i++) if (sm_last[i]) { span += (double)(sm_last[i] - sm_first[i]); nsm++; }
    if (cnt)
      fprintf(stderr, "[me_b200 item trace] %zu items on %d SMs, SM span %.0f clk; per item (clk): global atomic %.0f, decode+TMA issue %.0f, "
              "TMA issued -> first warp has the data %.0f, first warp -> last chunk handed out %.0f, last chunk -> last warp leaves %.0f, "
              "publish %.0f; items per SM %.1f => stage cycle budget %.0f clk per item per SM\n",
              cnt, nsm, span / nsm, d[0] / cnt, d[1] / cnt, d[2] / cnt, d[3] / cnt, d[4] / cnt, d[5] / cnt, (double)cnt / nsm,
              span / nsm / ((double)cnt / nsm));
    free(h);
    p.dbg = nullptr;
  }
#endif
  if (FORM == 4 && p.dbg) {   // measurements only
    unsigned long long h[8] = {0};
    cudaStreamSynchronize(s);
    cudaMemcpy(h, p.dbg, 64, cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    fprintf(stderr, "[me_b200] ssim form 4: periods %llu, drains with work %llu, drain iterations %llu, pops %llu, full evaluations %llu\n",
            h[4], h[0], h[1], h[2], h[3]);
  }
  if (e != cudaSuccess) *err = "tiled_search_kernel launch";
  else plan->outputs_enqueued = true;
  if (e == cudaSuccess && peer) plan->fused_launch = true;
  if (d_s) cudaFreeAsync(d_s, s);
  cudaFreeAsync(d_ctr, s);
  return e;
}

template <int WORDS, int BH, int NSUB, int FORM>
cudaError_t launch_shape(TiledPlan *plan, const Geom &g, const Frames &f, int npairs, const Out &o,
                         int by_begin, int by_count, cudaStream_t s, const char **err) {
  // PW: the frame width is not a multiple of the strip width, so the last strip holds a
  // partial-width block whose out-of-block reference bytes must be masked (FORM 1 only)
  if constexpr (FORM == 1) {
    if ((g.W % (4 * WORDS)) != 0)
      return launch_shape_pw<WORDS, BH, NSUB, 1, true>(plan, g, f, npairs, o, by_begin, by_count, s, err);
  }
  if constexpr (FORM >= 2) {
    // the energy table only knows full-width blocks, and its tile must fit the stage twice
    if (g.W % (4 * WORDS / NSUB) == 0) {
      cudaError_t e = launch_shape_pw<WORDS, BH, NSUB, FORM, false>(plan, g, f, npairs, o, by_begin, by_count, s, err);
      if (e != cudaErrorInvalidConfiguration) return e;
      (void)cudaGetLastError();
    }
    return launch_shape<WORDS, BH, NSUB, 1>(plan, g, f, npairs, o, by_begin, by_count, s, err);
  } else {
    return launch_shape_pw<WORDS, BH, NSUB, FORM, false>(plan, g, f, npairs, o, by_begin, by_count, s, err);
  }
}

}  // namespace

// SSIM-cost search of full-height, full-width 16x16 block rows [by_begin, by_begin + by_count) on the tiled kernel
// (FORM 4).  cudaErrorInvalidConfiguration: this geometry does not fit (the caller keeps its own kernel).
cudaError_t launch_tiled_ssim16(const Geom &g, const Frames &f, int npairs, const Out &o, int by_begin, int by_count,
                                const int2 *table, int table_y_lo, int table_rows, size_t table_pair_stride,
                                const int2 *blk_stats, cudaStream_t s, unsigned long long *launches) {
  if (g.B != 16 || by_count <= 0 || (by_begin + by_count) * 16 > g.H) return cudaErrorInvalidConfiguration;
  // (TMA: the table's row pitch, W entries of 8 bytes, must be a multiple of 16 bytes)
  if (!tiled_supported(g, f.pitch, f.pair_stride, f.cur, f.ref) || g.R > 127 || ((uintptr_t)table & 15) || (g.W & 1))
    return cudaErrorInvalidConfiguration;
  TiledPlan pl;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl.sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = scratch_pool(&pl.pool);
  if (e != cudaSuccess) return e;
  const char *pt = getenv("ME_B200_PARTS");
  const char *ns = getenv("ME_B200_NS");
  pl.parts_target = pt ? atoi(pt) : 0;
  pl.ns_override = ns ? atoi(ns) : 0;
  pl.ssim_table = table;
  pl.ssim_y_lo = table_y_lo;
  pl.ssim_rows = table_rows;
  pl.ssim_pair_stride = table_pair_stride;
  pl.ssim_blk = blk_stats;
  const char *err = nullptr;
  e = launch_shape_pw<4, 16, 1, 4, false>(&pl, g, f, npairs, o, by_begin, by_count, s, &err);
  if (launches) *launches += pl.kernels_launched;
  if (e != cudaSuccess && e != cudaErrorInvalidConfiguration)
    fprintf(stderr, "[me_b200] launch_tiled_ssim16 (%dx%d, span %d, rows %d+%d): %s: %s\n", g.W, g.H, g.R, by_begin, by_count,
            err ? err : "?", cudaGetErrorString(e));
  return e;
}

cudaError_t launch_tiled(TiledPlan *plan, const Geom &g, const Frames &f, int npairs, const Out &o,
                         cudaStream_t s, const char **err) {
  // Block rows of full height, and (FORM 1) a bottom row of exactly half height, run tiled;
  // any other partial bottom row is a different block shape and runs on the generic kernel
  // as a one-row band.
  plan->outputs_enqueued = false;
  struct ArriveReset {   // the arriving-frame settings apply to this launch only
    TiledPlan *p;
    ~ArriveReset() { p->arrive_flag = nullptr; p->arrive_status = nullptr; p->arrive_ref_resident = false; }
  } arrive_reset{plan};
  const int full_rows = g.H / g.B;
  const int hrem = g.H % g.B;
  const int tiled_rows = full_rows + ((plan->form >= 1 && hrem == g.B / 2) ? 1 : 0);
  const int r0 = g.by_begin, r1 = g.by_begin + g.by_count;
  const int t1 = r1 < tiled_rows ? r1 : tiled_rows;
  cudaError_t e = cudaSuccess;
  if (t1 > r0) {
    // the energy-table pre-pass (two small launches) only pays off once there is enough work
    const bool table = plan->form == 2 && (!plan->arrive_flag || plan->arrive_ref_resident) &&
                       (long long)npairs * g.W * g.H >= (plan->form_env_forced ? 0 : 2000000LL);
    // 8x8 blocks with the energy table on full-width frames: block rows whose window is not clamped
    // vertically go through the pair kernel two at a time, the rest through the single-row kernel
    int pa = t1, pb = t1;   // block rows [pa, pb) run as pairs
    // Measured (profiles/README.md): +12 % at +-8, +6 % at +-64, +1..2 % at +-12 / +-32 on 4K frames, but -1..-3 %
    // on 1080p / CIF frames at +-12 -- so: 4K-class frames and the spans where it clearly wins
    // ... all of that BEFORE the single-row kernel got one unrolled copy per period kind (ME_SPLIT_PERIODS): since
    // then it beats the pair kernel everywhere (4K +-12 54.9 vs 49.7 %, +-32 74.6 vs 67.5 %, Foreman CIF 43.4 vs 37.9 %
    // of the peak), so the pair kernel is opt-in (ME_B200_PAIR=1; the tests keep running it)
    const bool pair_pays = false;
    const char *pe = getenv("ME_B200_PAIR");   // 0 / 1 force it off / on (measurements, tests)
    const bool want_pair = pe ? pe[0] == '1' : pair_pays;
    if (table && want_pair && g.B == 8 && g.W % 8 == 0 && plan->npeer == 0 && !plan->arrive_flag) {
      const int first = (g.R + 7) / 8;                        // y0 >= R
      const int last = (g.H - g.R - 16) / 8;                  // y0 + 16 + R <= H  (upper row of the last pair)
      pa = r0 > first ? r0 : first;
      const int top = (t1 - 2 < last ? t1 - 2 : last);        // largest admissible upper row
      pb = top >= pa ? pa + ((top - pa) / 2 + 1) * 2 : pa;
      if (pb - pa < 8) pa = pb = t1;                          // not worth a launch of its own
    }
    if (table && pb > pa) {
      if (pa > r0) e = launch_shape<4, 8, 2, 2>(plan, g, f, npairs, o, r0, pa - r0, s, err);
      if (e == cudaSuccess) {
        e = launch_shape_pw<4, 8, 2, 2, false, 2>(plan, g, f, npairs, o, pa, pb - pa, s, err);
        if (e == cudaErrorInvalidConfiguration) {   // the bigger stage does not fit: single rows
          (void)cudaGetLastError();
          e = launch_shape<4, 8, 2, 2>(plan, g, f, npairs, o, pa, pb - pa, s, err);
        }
      }
      if (e == cudaSuccess && t1 > pb) e = launch_shape<4, 8, 2, 2>(plan, g, f, npairs, o, pb, t1 - pb, s, err);
    } else if (table) {
      // 16x16: FORM 3 (table with the bias, no sum cur^2 in the tasks) unless the field also goes to peers
      // (that publish path belongs to FORM 2) or ME_B200_FORM16=2 asks for the plain table (A/B measurements)
      const char *f16 = getenv("ME_B200_FORM16");
      const bool form16_plain = f16 && f16[0] == '2';
      if (g.B == 16 && plan->npeer == 0 && !form16_plain)
        e = launch_shape<4, 16, 1, 3>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
      else if (g.B == 16) e = launch_shape<4, 16, 1, 2>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
      else e = launch_shape<4, 8, 2, 2>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
    } else if (plan->form >= 1) {
      if (g.B == 16) e = launch_shape<4, 16, 1, 1>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
      else e = launch_shape<4, 8, 2, 1>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
    } else {
      if (g.B == 16) e = launch_shape<4, 16, 1, 0>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
      else e = launch_shape<8, 8, 4, 0>(plan, g, f, npairs, o, r0, t1 - r0, s, err);
    }
    if (e != cudaSuccess) return e;
    if (plan->npeer > 0 && plan->fused_launch) {
      plan->fused_begin = r0;
      plan->fused_end = t1;
    }
  }
  if (r1 > tiled_rows) {
    Geom gb = g;
    gb.by_begin = r0 > tiled_rows ? r0 : tiled_rows;
    gb.by_count = r1 - gb.by_begin;
    e = launch_generic(gb, f, npairs, o, s);
    plan->kernels_launched++;
    if (e != cudaSuccess) *err = "launch_generic(partial bottom row)";
  }
  return e;
}

}  // namespace me
