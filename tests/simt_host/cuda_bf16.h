// TEST INFRASTRUCTURE — host stand-in for <cuda_bf16.h>: bfloat16 as the upper half of a float, round to nearest even.
#pragma once
#include <stdint.h>
#include <string.h>
struct __nv_bfloat16 { uint16_t bits; };
inline __nv_bfloat16 __float2bfloat16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  __nv_bfloat16 r;
  if ((u & 0x7fffffffu) > 0x7f800000u) { r.bits = (uint16_t)((u >> 16) | 0x40); return r; }
  u += 0x7fffu + ((u >> 16) & 1u);
  r.bits = (uint16_t)(u >> 16);
  return r;
}
inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = (uint32_t)h.bits << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline unsigned short __bfloat16_as_ushort(__nv_bfloat16 h) { return h.bits; }
