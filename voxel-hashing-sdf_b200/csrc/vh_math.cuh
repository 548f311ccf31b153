// vh_math.cuh — IEEE binary32 arithmetic for the device, spelled so that nvcc can never contract it.
//
// Parity with the reference is defined against its sequential IEEE meaning (oracle/vh_oracle.c, built with
// -ffp-contract=off). The reference's own CUDA build used -use_fast_math (div.approx, fma contraction,
// FTZ; CMakeLists.txt:13); knife-edge results (roundf of a projection, DDA tie-breaks) flip between the two,
// so this engine computes every value that feeds a decision or a stored voxel with round-to-nearest
// single operations in the reference's expression order (/root/reference/src/tsdf.cu:67-116, :599-751).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vh {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float i2f(int a) { return __int2float_rn(a); }

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// a*b + c*d + e*f summed left to right, as `c2w[0]*t0 + c2w[4]*t1 + c2w[8]*t2` parses (tsdf.cu:86-92)
__device__ __forceinline__ float dot3_lr(float a, float b, float c, float d, float e, float f) {
  return fadd(fadd(fmul(a, b), fmul(c, d)), fmul(e, f));
}

struct Float3 { float x, y, z; };

// world -> camera: Rt (p - t), c2w row-major (base2cam, tsdf.cu:82-93)
__device__ __forceinline__ Float3 world_to_cam(const float* __restrict__ c2w, float px, float py, float pz) {
  const float t0 = fsub(px, c2w[3]), t1 = fsub(py, c2w[7]), t2 = fsub(pz, c2w[11]);
  Float3 c;
  c.x = dot3_lr(c2w[0], t0, c2w[4], t1, c2w[8], t2);
  c.y = dot3_lr(c2w[1], t0, c2w[5], t1, c2w[9], t2);
  c.z = dot3_lr(c2w[2], t0, c2w[6], t1, c2w[10], t2);
  return c;
}

// pixel + z-depth -> world (frame2cam + cam2base, tsdf.cu:67-73, :96-101): ((u-cx)*z)/fx, then R p + t left to right
__device__ __forceinline__ Float3 pixel_to_world(const float* __restrict__ c2w, float fx, float fy, float cx, float cy, int px, int py, float z) {
  const float x = fdiv(fmul(fsub(i2f(px), cx), z), fx);
  const float y = fdiv(fmul(fsub(i2f(py), cy), z), fy);
  Float3 w;
  w.x = fadd(fadd(fadd(fmul(x, c2w[0]), fmul(y, c2w[1])), fmul(z, c2w[2])), c2w[3]);
  w.y = fadd(fadd(fadd(fmul(x, c2w[4]), fmul(y, c2w[5])), fmul(z, c2w[6])), c2w[7]);
  w.z = fadd(fadd(fadd(fmul(x, c2w[8]), fmul(y, c2w[9])), fmul(z, c2w[10])), c2w[11]);
  return w;
}

}  // namespace vh
