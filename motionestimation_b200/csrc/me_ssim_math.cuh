// me_ssim_math.cuh -- the float arithmetic of the SSIM cost (src/common/ssim.c:44-60), shared by the SSIM kernels
// of me_ssim.cu and the SSIM formulation of the tiled search kernel (me_tiled.cu, FORM 4).  Every rounding
// operation is an explicit round-to-nearest intrinsic: nothing may be contracted into an FMA or reordered,
// because the reference build (gcc, x86-64, no FMA) rounds after every float operation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace me {

// ssim.c:47 -- double literals narrowed to float by the declaration
__device__ __forceinline__ float kC1() { return (float)0.01; }
__device__ __forceinline__ float kC2() { return (float)0.09; }
__device__ __forceinline__ float kC3() { return (float)0.045; }

// ssim.c:55-58 from the statistics of the two rectangles and the cross term (ssim.c:54)
__device__ __forceinline__ float ssim_from_stats(float mr, float sr, float mc, float sc, float cross) {
  const float lum = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.0f, mr), mc), kC1()),
                              __fadd_rn(__fadd_rn(__fmul_rn(mr, mr), __fmul_rn(mc, mc)), kC1()));
  const float con = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.0f, sr), sc), kC2()),
                              __fadd_rn(__fadd_rn(__fmul_rn(sr, sr), __fmul_rn(sc, sc)), kC2()));
  const float str = __fdiv_rn(__fadd_rn(cross, kC3()), __fadd_rn(__fmul_rn(sr, sc), kC3()));
  return __fmul_rn(__fmul_rn(lum, con), str);
}

// A finished candidate can only matter if its score can still reach the best score seen so far.
// score = fl(fl(L*C)*S) with L, C <= 1 up to rounding (at most 1 + 6.1u each, u = 2^-24), hence
// score <= (num/den) * (1 + 15.5u) for S = num/den > 0.  The test
//     fl(num * (1 + 2^-19)) < fl(thr * den)
// therefore proves score < thr (strictly: ties are never pruned, they are decided by the visit
// index) without a division; with thr = 0 it rejects exactly the candidates with num < 0, whose
// score cannot be above 0 (ssim.c:88,101).
__device__ __forceinline__ float kPruneMargin() { return 1.0000019073486328125f; }  // 1 + 2^-19

}  // namespace me
