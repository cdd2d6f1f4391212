// me_peak.cu -- integer-pipe microbenchmarks.  They define the roofline
// denominator for the search kernels: how many lane-
// instructions per second the SMs retire for IDP.4A.U8.U8, VABSDIFF4.U8 and for
// the VABSDIFF4+IDP.4A pair that scores four pixels.  Each kernel keeps 8
// independent dependency chains per thread so the pipes, not latency, bound it.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "me_b200.h"

namespace {

constexpr int kChains = 8;
constexpr int kUnroll = 16;  // ops per chain per loop trip

template <int WHICH>
__device__ __forceinline__ void step(uint32_t (&a)[kChains], uint32_t (&b)[kChains], uint32_t k,
                                     const uint32_t *lds, uint32_t &ldsacc, int u) {
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    if (WHICH == ME_PEAK_IDP4A) {
      a[c] = __dp4a(b[c], k, a[c]);
    } else if (WHICH == ME_PEAK_VABSDIFF4) {
      a[c] = __vabsdiffu4(a[c], b[c]);
    } else if (WHICH == ME_PEAK_SSD_PAIR || WHICH == ME_PEAK_SSD_PAIR_LDS) {
      uint32_t d = __vabsdiffu4(a[c], b[c]);  // chain-dependent, so nothing is hoisted
      a[c] = __dp4a(d, d, a[c]);
    } else if (WHICH == ME_PEAK_IADD3) {
      a[c] = a[c] + b[c] + k;
    } else if (WHICH == ME_PEAK_LOP3) {
      // explicit three-input LOP3s whose chain cannot be folded algebraically (round 1 measured an
      // impossible 928 lanes/clk/SM here: the compiler had collapsed `(a & b) ^ k` repeated 16 times):
      // majority and xor3 alternate, the second operand rotates over the chains
      if (u & 1)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[(c + 1) % kChains]), "r"(k));
      else
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(a[c]) : "r"(b[c]), "r"(k));
    } else if (WHICH == ME_PEAK_IMAD) {
      a[c] = a[c] * b[c] + k;
    } else if (WHICH == ME_PEAK_IDP4A_IADD3) {
      a[c] = __dp4a(b[c], k, a[c]);
      b[c] = b[c] + k + c;
    } else if (WHICH == ME_PEAK_VIMNMX) {
      a[c] = min(a[c], b[c]);
      a[c] = max(a[c], k);
    }
  }
  if (WHICH == ME_PEAK_SSD_PAIR_LDS) {
    if ((u & 1) == 0) ldsacc += lds[(threadIdx.x + u * 32 + ldsacc) & 1023];
  }
}

template <int WHICH>
__global__ void __launch_bounds__(512) peak_kernel(uint32_t *out, int iters, uint32_t seed,
                                                   unsigned long long *clk) {
  __shared__ uint32_t lds[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) lds[i] = (i * 2654435761u) >> 31;
  __syncthreads();
  uint32_t a[kChains], b[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    a[c] = seed * (c + 1) + threadIdx.x;
    b[c] = seed ^ (0x9e3779b9u * (c + 3)) ^ threadIdx.x;
  }
  uint32_t k = seed | 0x01010101u;
  uint32_t ldsacc = 0;
  unsigned long long t0 = 0, g0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    t0 = clock64();
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < kUnroll; u++) step<WHICH>(a, b, k, lds, ldsacc, u);
    k += 0x00010001u;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t1 = clock64(), g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    clk[0] = t1 - t0;
    clk[1] = g1 - g0;
  }
  uint32_t r = ldsacc;
#pragma unroll
  for (int c = 0; c < kChains; c++) r ^= a[c] + b[c];
  if (r == 0x12345679u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// Replica of the search kernel's hot loop without memory traffic or branches: 64 "current"
// registers, 16 rotating accumulators, 4 "reference" registers refreshed by one LOP3 each per
// row step, 64 IDP.4A per step in the same register pattern.  Measures what the FMA pipe
// sustains for that operand pattern at the search kernel's occupancy.
// VARIANT bit 0: two warp-uniform branches per row step (as the ramp skipping of the real loop);
//         bit 1: 5 LDS.32 + 4 funnel shifts per row step (the real loop's row fetch);
//         bit 2: start the warps of one scheduler at different loop phases.
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) replica_kernel(uint32_t *out, int iters, uint32_t seed,
                                                         unsigned long long *clk, int flags) {
  __shared__ uint32_t rows[64 * 64];
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) rows[i] = seed * (i + 1);
  __syncthreads();
  uint32_t cur[16][4], acc[16], ref[4];
#pragma unroll
  for (int r = 0; r < 16; r++)
#pragma unroll
    for (int w = 0; w < 4; w++) cur[r][w] = seed * (r * 4 + w + 1) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < 16; r++) acc[r] = 0;
#pragma unroll
  for (int w = 0; w < 4; w++) ref[w] = seed ^ (threadIdx.x * (w + 7));
  unsigned long long t0 = 0, g0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    t0 = clock64();
  }
  if (VARIANT & 4) {
    const long long until = clock64() + (long long)(threadIdx.x >> 7) * 137;
    while (clock64() < until) {
    }
  }
  uint32_t best = 0xffffffffu;
  const uint32_t *rowp = rows + (threadIdx.x & 31);
  const uint32_t shift = 8u * (threadIdx.x & 3);
  const bool fa = (flags & 1) != 0, fb = (flags & 2) != 0;   // both true at run time
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int s_ = 0; s_ < 16; s_++) {
      if (VARIANT & 2) {
        uint32_t raw[5];
#pragma unroll
        for (int w = 0; w < 5; w++) raw[w] = rowp[s_ * 64 + w];
#pragma unroll
        for (int w = 0; w < 4; w++) ref[w] = __funnelshift_r(raw[w], raw[w + 1], shift) ^ best;
      } else {
#pragma unroll
        for (int w = 0; w < 4; w++) ref[w] = (ref[w] ^ best) + 0x01010101u * (w + 1);
      }
      auto group = [&](const int r) {
        const int slot = (s_ - r + 16) % 16;
        uint32_t a = r == 0 ? 0u : acc[slot];
#pragma unroll
        for (int w = 0; w < 4; w++) a = __dp4a(cur[r][w], ref[w], a);
        acc[slot] = a;
        if (r == 15) best = min(best, (a << 8) + (uint32_t)s_);
      };
      if (VARIANT & 1) {
        if (fa) {
#pragma unroll
          for (int r = 0; r < s_; r++) group(r);
        }
        if (fb) {
#pragma unroll
          for (int r = s_ + 1; r < 16; r++) group(r);
        }
        group(s_);
      } else {
#pragma unroll
        for (int r = 0; r < 16; r++) group(r);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t1 = clock64(), g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    clk[0] = t1 - t0;
    clk[1] = g1 - g0;
  }
  if (best == 0x12345679u) out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

template <int VARIANT>
double run_replica(int iters, double *mhz) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0.0;
  const char *et = getenv("ME_PEAK_THREADS");
  const int ctas = sms, threads = et ? atoi(et) : 512;
  uint32_t *out = nullptr;
  unsigned long long *clk = nullptr;
  if (cudaMalloc(&out, (size_t)ctas * threads * 4) != cudaSuccess) return 0.0;
  if (cudaMalloc(&clk, 16) != cudaSuccess) { cudaFree(out); return 0.0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  replica_kernel<VARIANT><<<ctas, threads>>>(out, iters / 8 + 1, 12345u, clk, 3);
  cudaEventRecord(e0);
  replica_kernel<VARIANT><<<ctas, threads>>>(out, iters, 12345u, clk, 3);
  cudaEventRecord(e1);
  double rate = 0.0;
  if (cudaEventSynchronize(e1) == cudaSuccess) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2] = {0, 0};
    cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
    if (mhz) *mhz = h[1] ? (double)h[0] / (double)h[1] * 1000.0 : 0.0;
    rate = (double)ctas * threads * (double)iters * 16 * 64 / (ms * 1e-3);  // IDP.4A lane-instructions / s
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaFree(clk);
  return rate;
}

template <int WHICH>
double run_peak(int iters, double *mhz) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0.0;
  // default: 8 CTAs x 256 threads per SM (64 warps/SM); env overrides let the bench probe the
  // search kernel's own occupancy (1 CTA x 512 threads)
  const char *ec = getenv("ME_PEAK_CTAS_PER_SM"), *et = getenv("ME_PEAK_THREADS");
  const int ctas = sms * (ec ? atoi(ec) : 8), threads = et ? atoi(et) : 256;
  uint32_t *out = nullptr;
  unsigned long long *clk = nullptr;
  if (cudaMalloc(&out, (size_t)ctas * threads * 4) != cudaSuccess) return 0.0;
  if (cudaMalloc(&clk, 16) != cudaSuccess) { cudaFree(out); return 0.0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  peak_kernel<WHICH><<<ctas, threads>>>(out, iters / 8 + 1, 12345u, clk);  // warm-up
  cudaEventRecord(e0);
  peak_kernel<WHICH><<<ctas, threads>>>(out, iters, 12345u, clk);
  cudaEventRecord(e1);
  double rate = 0.0;
  if (cudaEventSynchronize(e1) == cudaSuccess) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[2] = {0, 0};
    cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
    if (mhz) *mhz = h[1] ? (double)h[0] / (double)h[1] * 1000.0 : 0.0;
    // counted op per chain step: the named op (pair = 1 VABSDIFF4 + 1 IDP.4A = 2)
    double per = (WHICH == ME_PEAK_SSD_PAIR || WHICH == ME_PEAK_SSD_PAIR_LDS || WHICH == ME_PEAK_VIMNMX ||
                  WHICH == ME_PEAK_IDP4A_IADD3) ? 2.0 : 1.0;
    double ops = (double)ctas * threads * (double)iters * kUnroll * kChains * per;
    rate = ops / (ms * 1e-3);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  cudaFree(clk);
  return rate;
}

}  // namespace

extern "C" double me_b200_int_peak(int device, int which, int iters, double *sm_clock_mhz) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return 0.0;
  if (cudaSetDevice(device) != cudaSuccess) return 0.0;
  if (iters < 1) iters = 1;
  switch (which) {
    case ME_PEAK_IDP4A: return run_peak<ME_PEAK_IDP4A>(iters, sm_clock_mhz);
    case ME_PEAK_VABSDIFF4: return run_peak<ME_PEAK_VABSDIFF4>(iters, sm_clock_mhz);
    case ME_PEAK_SSD_PAIR: return run_peak<ME_PEAK_SSD_PAIR>(iters, sm_clock_mhz);
    case ME_PEAK_IADD3: return run_peak<ME_PEAK_IADD3>(iters, sm_clock_mhz);
    case ME_PEAK_LOP3: return run_peak<ME_PEAK_LOP3>(iters, sm_clock_mhz);
    case ME_PEAK_IMAD: return run_peak<ME_PEAK_IMAD>(iters, sm_clock_mhz);
    case ME_PEAK_VIMNMX: return run_peak<ME_PEAK_VIMNMX>(iters, sm_clock_mhz);
    case ME_PEAK_SSD_PAIR_LDS: return run_peak<ME_PEAK_SSD_PAIR_LDS>(iters, sm_clock_mhz);
    case ME_PEAK_IDP4A_IADD3: return run_peak<ME_PEAK_IDP4A_IADD3>(iters, sm_clock_mhz);
    case ME_PEAK_LOOP_REPLICA: {
      const char *v = getenv("ME_PEAK_REPLICA_VARIANT");
      switch (v ? atoi(v) : 0) {
        case 1: return run_replica<1>(iters / 16 + 1, sm_clock_mhz);
        case 2: return run_replica<2>(iters / 16 + 1, sm_clock_mhz);
        case 3: return run_replica<3>(iters / 16 + 1, sm_clock_mhz);
        case 7: return run_replica<7>(iters / 16 + 1, sm_clock_mhz);
        case 4: return run_replica<4>(iters / 16 + 1, sm_clock_mhz);
        default: return run_replica<0>(iters / 16 + 1, sm_clock_mhz);
      }
    }
    default: return 0.0;
  }
}
