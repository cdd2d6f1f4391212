// me_tma.cuh -- TMA / mbarrier PTX helpers shared by the sm_100a kernels (me_tiled.cu, me_direct.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace me {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 3-D tile load (x, y, pair), bytes signalled on `bar`; dst is a shared-window address
__device__ __forceinline__ void tma_load_3d_s(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
  tma_load_3d_s(smem_u32(dst), map, bar, x, y, z);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked directly)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// u8 frames of a batch as a 3-D tensor (x, y, pair) with a box of box_w x box_h x 1 BYTES; out-of-frame
// coordinates (negative or beyond W / H) are legal and zero-filled.  A box dimension is limited to 256
// elements, so wide boxes (up to 1024 bytes per row) are described in 4-byte elements (words = true:
// W, box_w and every x coordinate must then be multiples of 4, and x is given in words).
inline bool encode_frames_map(CUtensorMap *map, const void *base, int W, int H, int npairs, size_t pitch,
                              size_t pair_stride, int box_w, int box_h, bool words = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const int eb = words ? 4 : 1;
  const cuuint64_t dims[3] = {(cuuint64_t)(W / eb), (cuuint64_t)H, (cuuint64_t)npairs};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(npairs > 1 ? pair_stride : pitch * (size_t)H)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t box[3] = {(cuuint32_t)(box_w / eb), (cuuint32_t)box_h, 1};
  return enc(map, words ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides,
             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace me
