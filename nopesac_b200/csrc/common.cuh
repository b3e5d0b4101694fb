// Shared device helpers + host-side error plumbing for libnopesac_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nopesac_b200.h"

// ---------------------------------------------------------------------------------------------
// host: thread-local error string, argument / launch checks
// ---------------------------------------------------------------------------------------------
void nsac_set_error(const char* fmt, ...);

#define NSAC_REQUIRE(cond, ...)                        \
  do {                                                 \
    if (!(cond)) {                                     \
      nsac_set_error(__VA_ARGS__);                     \
      return NSAC_ERR_ARG;                             \
    }                                                  \
  } while (0)

#define NSAC_CHECK_LAUNCH(what)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      nsac_set_error("%s: %s", what, cudaGetErrorString(e__));                      \
      return NSAC_ERR_LAUNCH;                                                       \
    }                                                                               \
  } while (0)

#define NSAC_CUDA(call)                                                             \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      nsac_set_error("%s: %s", #call, cudaGetErrorString(e__));                     \
      return NSAC_ERR_LAUNCH;                                                       \
    }                                                                               \
  } while (0)

static inline int nsac_cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#define NSAC_FULL_MASK 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSAC_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(NSAC_FULL_MASK, v, o));
  return v;
}

struct Mat3 {
  float m[9];
};

// (w,x,y,z) -> R, element formulas of camera_head.py:1148-1173 (same association order).
__device__ __forceinline__ Mat3 quat_to_rot(float w, float x, float y, float z) {
  Mat3 R;
  R.m[0] = 1.f - 2.f * y * y - 2.f * z * z;
  R.m[1] = 2.f * x * y - 2.f * w * z;
  R.m[2] = 2.f * x * z + 2.f * w * y;
  R.m[3] = 2.f * x * y + 2.f * w * z;
  R.m[4] = 1.f - 2.f * x * x - 2.f * z * z;
  R.m[5] = 2.f * y * z - 2.f * w * x;
  R.m[6] = 2.f * x * z - 2.f * w * y;
  R.m[7] = 2.f * y * z + 2.f * w * x;
  R.m[8] = 1.f - 2.f * x * x - 2.f * y * y;
  return R;
}

// Reference plane warp (camera_head.py:1446-1453): end = R (p*flip) + t ; b = end - t ;
// pi = (end.b / (|b| + 1e-5)^2) b.   Kept in the reference's operation order because its output
// feeds discrete decisions (sig_seq, pruning thresholds).
__device__ __forceinline__ void warp_plane(const Mat3& R, float tx, float ty, float tz, float px,
                                           float py, float pz, float& ox, float& oy, float& oz) {
  const float fx = px, fy = -py, fz = -pz;
  const float ex = R.m[0] * fx + R.m[1] * fy + R.m[2] * fz + tx;
  const float ey = R.m[3] * fx + R.m[4] * fy + R.m[5] * fz + ty;
  const float ez = R.m[6] * fx + R.m[7] * fy + R.m[8] * fz + tz;
  const float bx = ex - tx, by = ey - ty, bz = ez - tz;
  const float ab = ex * bx + ey * by + ez * bz;
  const float nb = sqrtf(bx * bx + by * by + bz * bz) + 1e-5f;
  const float k = ab / (nb * nb);
  ox = k * bx;
  oy = k * by;
  oz = k * bz;
}

// F.normalize(v, eps=1e-12): v / max(|v|, 1e-12); also returns |v|.
__device__ __forceinline__ float normalize3(float& x, float& y, float& z) {
  const float n = sqrtf(x * x + y * y + z * z);
  const float d = fmaxf(n, 1e-12f);
  x /= d;
  y /= d;
  z /= d;
  return n;
}
