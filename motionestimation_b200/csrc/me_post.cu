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

// Batched + vectorised variant (frame width a multiple of 16, 16-byte aligned frames / planes, block
// size a multiple of 4): a thread owns 16 consecutive pixels of one row of one pair -- one 16-byte
// load of the current and of the reference row, the motion-compensated bytes as up to five aligned
// words around ref[(y+mvy)][x+mvx] per 4-pixel word (served by L1/L2: the reference frame is being
// streamed anyway) byte-aligned with funnel shifts, |a-b| with VABSDIFF4, the squared error with
// IDP.4A, five 16-byte stores.  Algorithmic HBM bytes: 2 read + 5 written per pixel.
__device__ __forceinline__ uint32_t umax4(uint32_t v) {
  return max(max(v & 0xffu, (v >> 8) & 0xffu), max((v >> 16) & 0xffu, v >> 24));
}

__global__ void __launch_bounds__(256)
post_batch_kernel(Geom g, const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref, size_t pitch,
                  size_t pair_stride, const int32_t *__restrict__ mvx, const int32_t *__restrict__ mvy,
                  uint8_t *__restrict__ out5, size_t out_pair_stride, unsigned long long *sq_err, uint32_t *mx) {
  const int pair = blockIdx.y;
  const uint8_t *c0 = cur + (size_t)pair * pair_stride, *r0 = ref + (size_t)pair * pair_stride;
  const int32_t *vx = mvx + (size_t)pair * g.nbx * g.nby, *vy = mvy + (size_t)pair * g.nbx * g.nby;
  uint8_t *o = out5 + (size_t)pair * out_pair_stride;
  const size_t plane = (size_t)g.W * g.H;
  const int gpr = g.W >> 4;                     // 16-pixel groups per row
  const int total = gpr * g.H;
  unsigned long long sq = 0;
  uint32_t peak = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / gpr, x = (i - y * gpr) << 4;
    const uint4 cv = *reinterpret_cast<const uint4 *>(c0 + (size_t)y * pitch + x);
    const uint4 rv = *reinterpret_cast<const uint4 *>(r0 + (size_t)y * pitch + x);
    const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t mw[4], d0[4], d1[4];
    const int brow = (y / g.B) * g.nbx;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int xq = x + 4 * q;
      const int bi = brow + xq / g.B;             // B is a multiple of 4: the whole word lies in one block
      const int sx = xq + vx[bi], sy = y + vy[bi];
      uint32_t m = 0;
      // full search never leaves the frame; the test only guards foreign MV fields
      // (the reference leaves such pixels unwritten, utils.c:122)
      if (sx >= 0 && sy >= 0 && sx + 4 <= g.W && sy < g.H) {
        const uint8_t *a = r0 + (size_t)sy * pitch + sx;
        const uint32_t *aw = reinterpret_cast<const uint32_t *>(a - ((uintptr_t)a & 3));
        const uint32_t sh = 8u * (uint32_t)((uintptr_t)a & 3);
        const uint32_t lo = aw[0], hi = sh ? aw[1] : 0u;
        m = __funnelshift_r(lo, hi, sh);
      } else {
        for (int b = 0; b < 4; b++) {
          const int px = sx + b;
          if (px >= 0 && sy >= 0 && px < g.W && sy < g.H) m |= (uint32_t)r0[(size_t)sy * pitch + px] << (8 * b);
        }
      }
      mw[q] = m;
      d0[q] = __vabsdiffu4(rw[q], cw[q]);         // |ref - cur|  (utils.c:94-100)
      d1[q] = __vabsdiffu4(m, cw[q]);             // |mc - cur|
      sq += __dp4a(d1[q], d1[q], 0u);
      peak = max(peak, max(umax4(m), umax4(cw[q])));
    }
    uint8_t *dst = o + (size_t)y * g.W + x;
    *reinterpret_cast<uint4 *>(dst) = rv;
    *reinterpret_cast<uint4 *>(dst + plane) = cv;
    *reinterpret_cast<uint4 *>(dst + 2 * plane) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
    *reinterpret_cast<uint4 *>(dst + 3 * plane) = make_uint4(d0[0], d0[1], d0[2], d0[3]);
    *reinterpret_cast<uint4 *>(dst + 4 * plane) = make_uint4(d1[0], d1[1], d1[2], d1[3]);
  }
  for (int off = 16; off; off >>= 1) {
    sq += __shfl_down_sync(0xffffffffu, sq, off);
    peak = max(peak, __shfl_down_sync(0xffffffffu, peak, off));
  }
  __shared__ unsigned long long s_sq[8];
  __shared__ uint32_t s_pk[8];
  if ((threadIdx.x & 31) == 0) {
    s_sq[threadIdx.x >> 5] = sq;
    s_pk[threadIdx.x >> 5] = peak;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) {
      sq += s_sq[k];
      peak = max(peak, s_pk[k]);
    }
    if (sq_err) atomicAdd(sq_err + pair, sq);
    if (mx) atomicMax(mx + pair, peak);
  }
}

}  // namespace

cudaError_t launch_postprocess_batch(const Geom &g, const uint8_t *cur, const uint8_t *ref, size_t pitch,
                                     size_t pair_stride, int npairs, const int32_t *mvx, const int32_t *mvy,
                                     uint8_t *out5, size_t out_pair_stride, unsigned long long *sq_err, uint32_t *mx,
                                     cudaStream_t s) {
  cudaError_t e;
  if (sq_err && (e = cudaMemsetAsync(sq_err, 0, sizeof(unsigned long long) * (size_t)npairs, s)) != cudaSuccess) return e;
  if (mx && (e = cudaMemsetAsync(mx, 0, sizeof(uint32_t) * (size_t)npairs, s)) != cudaSuccess) return e;
  const bool vec = (g.W & 15) == 0 && (g.B & 3) == 0 && (pitch & 15) == 0 && (pair_stride & 15) == 0 &&
                   (out_pair_stride & 15) == 0 &&
                   ((((uintptr_t)cur) | ((uintptr_t)ref) | ((uintptr_t)out5)) & 15) == 0;
  if (vec) {
    const int total = (g.W >> 4) * g.H;
    int ctas = (total + 255) / 256;
    // enough CTAs to fill the machine a few times over, few enough that the per-CTA reduction stays cheap
    const int cap = (148 * 8 * 4 + npairs - 1) / npairs;
    if (ctas > cap) ctas = cap < 1 ? 1 : cap;
    for (int done = 0; done < npairs; done += 65535) {
      const int n = npairs - done > 65535 ? 65535 : npairs - done;
      post_batch_kernel<<<dim3((unsigned)ctas, (unsigned)n), 256, 0, s>>>(
          g, cur + (size_t)done * pair_stride, ref + (size_t)done * pair_stride, pitch, pair_stride,
          mvx + (size_t)done * g.nbx * g.nby, mvy + (size_t)done * g.nbx * g.nby, out5 + (size_t)done * out_pair_stride,
          out_pair_stride, sq_err ? sq_err + done : nullptr, mx ? mx + done : nullptr);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  // any other layout: the per-pixel kernel, pair by pair
  const long long total = (long long)g.W * g.H;
  long long ctas = (total + 255) / 256;
  if (ctas > 148 * 16) ctas = 148 * 16;
  for (int p = 0; p < npairs; p++) {
    post_kernel<<<(unsigned)ctas, 256, 0, s>>>(g, cur + (size_t)p * pair_stride, ref + (size_t)p * pair_stride, pitch,
                                               mvx + (size_t)p * g.nbx * g.nby, mvy + (size_t)p * g.nbx * g.nby,
                                               out5 + (size_t)p * out_pair_stride, sq_err ? sq_err + p : nullptr,
                                               mx ? mx + p : nullptr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

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
