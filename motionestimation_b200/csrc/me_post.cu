// me_post.cu -- post-search stage on the device.
// One pass builds the reference's 5 stacked output planes
//   ref, cur, motion-compensated, |ref-cur|, |mc-cur|        (main.c:160-168)
// from the motion-vector field (gather, utils.c:110-129; |a-b|, utils.c:94-100)
// and reduces the two integers imagePSNR needs (utils.c:146-156): the sum of
// squared (mc-cur) and the largest pixel of mc and cur.  HBM-bound: 2 bytes read
// (+1 gathered) and 5 bytes written per pixel, one pixel per thread so every
// warp access is a contiguous 32-byte sector.
#include "me_device.cuh"

namespace me {

namespace {

__global__ void __launch_bounds__(256)
post_kernel(Geom g, const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref, size_t pitch,
            const int32_t *__restrict__ mvx, const int32_t *__restrict__ mvy,
            uint8_t *__restrict__ out5, unsigned long long *sq_err, uint32_t *mx) {
  const size_t plane = (size_t)g.W * g.H;
  unsigned long long sq = 0;
  uint32_t peak = 0;
  const long long total = (long long)g.W * g.H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / g.W), x = (int)(i - (long long)y * g.W);
    const int bi = (y / g.B) * g.nbx + x / g.B;
    const int sx = x + mvx[bi], sy = y + mvy[bi];
    const uint32_t c = cur[(size_t)y * pitch + x];
    const uint32_t r = ref[(size_t)y * pitch + x];
    // full search never leaves the frame; the clamp only guards foreign MV fields
    // (the reference leaves such pixels unwritten, utils.c:122)
    uint32_t m = 0;
    if (sx >= 0 && sy >= 0 && sx < g.W && sy < g.H) m = ref[(size_t)sy * pitch + sx];
    const uint32_t d_rc = r > c ? r - c : c - r;
    const uint32_t d_mc = m > c ? m - c : c - m;
    out5[i] = (uint8_t)r;
    out5[plane + i] = (uint8_t)c;
    out5[2 * plane + i] = (uint8_t)m;
    out5[3 * plane + i] = (uint8_t)d_rc;
    out5[4 * plane + i] = (uint8_t)d_mc;
    sq += (unsigned long long)(d_mc * d_mc);
    peak = max(peak, max(m, c));
  }
  for (int off = 16; off; off >>= 1) {
    sq += __shfl_down_sync(0xffffffffu, sq, off);
    peak = max(peak, __shfl_down_sync(0xffffffffu, peak, off));
  }
  if ((threadIdx.x & 31) == 0) {
    if (sq_err) atomicAdd(sq_err, sq);
    if (mx) atomicMax(mx, peak);
  }
}

}  // namespace

cudaError_t launch_postprocess(const Geom &g, const uint8_t *cur, const uint8_t *ref, size_t pitch,
                               const int32_t *mvx, const int32_t *mvy, uint8_t *out5,
                               unsigned long long *sq_err, uint32_t *mx, cudaStream_t s) {
  cudaError_t e;
  if (sq_err && (e = cudaMemsetAsync(sq_err, 0, sizeof(unsigned long long), s)) != cudaSuccess) return e;
  if (mx && (e = cudaMemsetAsync(mx, 0, sizeof(uint32_t), s)) != cudaSuccess) return e;
  const long long total = (long long)g.W * g.H;
  long long ctas = (total + 255) / 256;
  if (ctas > 148 * 16) ctas = 148 * 16;
  post_kernel<<<(unsigned)ctas, 256, 0, s>>>(g, cur, ref, pitch, mvx, mvy, out5, sq_err, mx);
  return cudaGetLastError();
}

}  // namespace me
