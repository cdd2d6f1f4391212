// tma_probe.cu -- development probe: does a 3-D u8 TMA tile load accept byte-granular
// (unaligned, negative) coordinates and zero-fill out-of-frame bytes?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu
// run:   /tmp/tma_probe <x> <y> <boxw> <boxh>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bytes, uint8_t *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(d), "l"(&map), "r"(b), "r"(x), "r"(y), "r"(z) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }"
                 : "=r"(ok) : "r"(b) : "memory");
  }
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char **argv) {
  int x = argc > 1 ? atoi(argv[1]) : 0, y = argc > 2 ? atoi(argv[2]) : 0;
  int bw = argc > 3 ? atoi(argv[3]) : 32, bh = argc > 4 ? atoi(argv[4]) : 4;
  const int W = 64, H = 16, P = 2;
  uint8_t h[P * H * W];
  for (int p = 0; p < P; p++) for (int r = 0; r < H; r++) for (int c = 0; c < W; c++) h[(p * H + r) * W + c] = (uint8_t)(p * 100 + r * 4 + c % 4 + (c / 4) * 0 + c);
  uint8_t *d, *o;
  cudaMalloc(&d, sizeof h); cudaMalloc(&o, bw * bh);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap m;
  cuuint64_t dims[3] = {W, H, P}, strides[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d\n", (int)r);
  probe<<<1, 64, bw * bh>>>(m, x, y, 1, bw * bh, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("x=%d y=%d box=%dx%d: %s\n", x, y, bw, bh, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    uint8_t *res = (uint8_t *)malloc(bw * bh);
    cudaMemcpy(res, o, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < bh; rr++) for (int c = 0; c < bw; c++) {
      int gx = x + c, gy = y + rr;
      uint8_t want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0 : h[(1 * H + gy) * W + gx];
      if (res[rr * bw + c] != want) bad++;
    }
    printf("mismatches=%d first row: ", bad);
    for (int c = 0; c < 16; c++) printf("%d ", res[c]);
    printf("\n");
  }
  return 0;
}
