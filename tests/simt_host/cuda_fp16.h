// TEST INFRASTRUCTURE — host stand-in for <cuda_fp16.h>: IEEE binary16 through GCC's _Float16 (round to nearest even).
#pragma once
#include <string.h>
typedef _Float16 __half;
inline __half __float2half_rn(float x) { return (__half)x; }
inline float __half2float(__half h) { return (float)h; }
inline unsigned short __half_as_ushort(__half h) {
  unsigned short u;
  memcpy(&u, &h, 2);
  return u;
}
