// Glue kernels of the ResNet-50 backbone (detectron2 `build_resnet_backbone`, Base.yaml:2-12 — plain
// ResNet, STRIDE_IN_1X1 = False, FrozenBN) around the tensor-core engine: every convolution of the backbone is a GEMM on
// NHWC 16-bit hi/lo planes (1x1: nsac_gemm_split; 3x3 stride 1: nsac_conv3x3_split; 3x3 stride 2: nsac_im2col3x3_planes +
// nsac_gemm_split; FrozenBN folded into weights + bias, ReLU in the epilogue).  What is left for the CUDA cores, all HBM-bound
// byte movers:
//   stem_im2col_kernel     (x - PIXEL_MEAN) / PIXEL_STD fused with the im2col of the 7x7 / stride 2 / pad 3 stem convolution on
//                          the 3-channel NCHW image -> planes [N*Ho*Wo, 192] (K = 147 in (ky,kx,c) order, zero padded)
//   maxpool3x3s2_kernel    MaxPool2d(3, 2, 1) on NHWC fp32 -> fp32 and / or planes
//   subsample2_kernel      every second pixel of NHWC planes (the input of a stride-2 1x1 shortcut convolution)
//   add_relu_kernel        relu(a + b) of two NHWC fp32 maps (bottleneck output) -> fp32 + planes for the next block
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {
__device__ __forceinline__ void bb_split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == NSAC_SPLIT_F16) {
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
  }
}

constexpr int STEM_K = 147, STEM_KP = 192;

__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, int N, int H, int W, int Ho, int Wo, float m0, float m1, float m2, float s0,
                   float s1, float s2, int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const size_t total = (size_t)N * Ho * Wo * STEM_KP;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % STEM_KP);
    const size_t row = idx / STEM_KP;
    float v = 0.f;
    if (k < STEM_K) {
      const int c = k % 3, tap = k / 3, ky = tap / 7, kx = tap - ky * 7;
      const int xo = (int)(row % Wo), yo = (int)((row / Wo) % Ho), n = (int)(row / ((size_t)Wo * Ho));
      const int y = 2 * yo + ky - 3, x = 2 * xo + kx - 3;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float p = __ldg(img + (((size_t)n * 3 + c) * H + y) * W + x);
        v = (p - (c == 0 ? m0 : (c == 1 ? m1 : m2))) / (c == 0 ? s0 : (c == 1 ? s1 : s2));
      }
    }
    uint16_t h, l;
    bb_split16(v, fmt, h, l);
    hi[idx] = h;
    lo[idx] = l;
  }
}

// one thread = 4 channels of one output pixel
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, int fmt, float* __restrict__ out_f32,
                    uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  const int C4 = C >> 2;
  const size_t total = (size_t)N * Ho * Wo * C4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    const size_t pix = idx / C4;
    const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * yo + dy - 1;
      if (y < 0 || y >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * xo + dx - 1;
        if (xx < 0 || xx >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + y) * W + xx) * C + c4 * 4));
        m[0] = fmaxf(m[0], v.x); m[1] = fmaxf(m[1], v.y); m[2] = fmaxf(m[2], v.z); m[3] = fmaxf(m[3], v.w);
      }
    }
    const size_t o = pix * C + c4 * 4;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(m[0], m[1], m[2], m[3]);
    if (hi) {
      uint16_t h[4], l[4];
      for (int i = 0; i < 4; ++i) bb_split16(m[i], fmt, h[i], l[i]);
      *reinterpret_cast<uint2*>(hi + o) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      *reinterpret_cast<uint2*>(lo + o) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
  }
}

// one thread = 8 channels (16 bytes) of one output pixel, both planes
__global__ void __launch_bounds__(256)
subsample2_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int N, int H, int W, int C, int Ho, int Wo,
                  uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int C8 = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % C8);
    const size_t pix = idx / C8;
    const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    const size_t src = (((size_t)n * H + 2 * yo) * W + 2 * xo) * C + c8 * 8, dst = pix * C + c8 * 8;
    *reinterpret_cast<uint4*>(out_hi + dst) = __ldg(reinterpret_cast<const uint4*>(hi + src));
    *reinterpret_cast<uint4*>(out_lo + dst) = __ldg(reinterpret_cast<const uint4*>(lo + src));
  }
}

// one thread = 4 consecutive elements
__global__ void __launch_bounds__(256)
add_relu_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n4, int fmt, float* __restrict__ out_f32,
                uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n4; idx += (size_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + idx), y = __ldg(reinterpret_cast<const float4*>(b) + idx);
    const float r[4] = {fmaxf(x.x + y.x, 0.f), fmaxf(x.y + y.y, 0.f), fmaxf(x.z + y.z, 0.f), fmaxf(x.w + y.w, 0.f)};
    if (out_f32) reinterpret_cast<float4*>(out_f32)[idx] = make_float4(r[0], r[1], r[2], r[3]);
    if (hi) {
      uint16_t h[4], l[4];
      for (int i = 0; i < 4; ++i) bb_split16(r[i], fmt, h[i], l[i]);
      reinterpret_cast<uint2*>(hi)[idx] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
      reinterpret_cast<uint2*>(lo)[idx] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
  }
}

inline int grid_for(size_t total) {
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)148 * 16;          // a few waves of the 148 SMs, grid-stride beyond that
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}
}  // namespace

extern "C" int nsac_stem_im2col_planes(const float* img, int N, int H, int W, const float* mean3_host, const float* std3_host,
                                       int fmt, void* hi, void* lo, void* stream) {
  NSAC_REQUIRE(img && mean3_host && std3_host && hi && lo, "nsac_stem_im2col_planes: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 7 && W >= 7, "nsac_stem_im2col_planes: bad shape N=%d H=%d W=%d", N, H, W);
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_stem_im2col_planes: bad plane format %d", fmt);
  NSAC_REQUIRE(std3_host[0] != 0.f && std3_host[1] != 0.f && std3_host[2] != 0.f, "nsac_stem_im2col_planes: zero PIXEL_STD");
  if (N == 0) return NSAC_OK;
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  stem_im2col_kernel<<<grid_for((size_t)N * Ho * Wo * STEM_KP), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, N, H, W, Ho, Wo, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], fmt, static_cast<uint16_t*>(hi),
      static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_stem_im2col_planes");
  return NSAC_OK;
}

extern "C" int nsac_maxpool3x3s2_nhwc(const float* x, int N, int H, int W, int C, int fmt, float* out_f32, void* hi, void* lo,
                                      void* stream) {
  NSAC_REQUIRE(x && (out_f32 || (hi && lo)), "nsac_maxpool3x3s2_nhwc: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "nsac_maxpool3x3s2_nhwc: bad shape (C %% 4 == 0)");
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_maxpool3x3s2_nhwc: bad plane format %d", fmt);
  if (N == 0) return NSAC_OK;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3x3s2_kernel<<<grid_for((size_t)N * Ho * Wo * (C / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, N, H, W, C, Ho, Wo, fmt, out_f32, static_cast<uint16_t*>(hi), static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_maxpool3x3s2_nhwc");
  return NSAC_OK;
}

extern "C" int nsac_subsample2_planes(const void* hi, const void* lo, int N, int H, int W, int C, void* out_hi, void* out_lo,
                                      void* stream) {
  NSAC_REQUIRE(hi && lo && out_hi && out_lo, "nsac_subsample2_planes: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, "nsac_subsample2_planes: bad shape (C %% 8 == 0)");
  if (N == 0) return NSAC_OK;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  subsample2_kernel<<<grid_for((size_t)N * Ho * Wo * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(hi), static_cast<const uint16_t*>(lo), N, H, W, C, Ho, Wo, static_cast<uint16_t*>(out_hi),
      static_cast<uint16_t*>(out_lo));
  NSAC_CHECK_LAUNCH("nsac_subsample2_planes");
  return NSAC_OK;
}

extern "C" int nsac_add_relu_nhwc(const float* a, const float* b, size_t count, int fmt, float* out_f32, void* hi, void* lo,
                                  void* stream) {
  NSAC_REQUIRE(a && b && (out_f32 || (hi && lo)), "nsac_add_relu_nhwc: null pointer");
  NSAC_REQUIRE(count % 4 == 0, "nsac_add_relu_nhwc: element count must be a multiple of 4");
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_add_relu_nhwc: bad plane format %d", fmt);
  if (count == 0) return NSAC_OK;
  add_relu_kernel<<<grid_for(count / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, count / 4, fmt, out_f32,
                                                                                      static_cast<uint16_t*>(hi),
                                                                                      static_cast<uint16_t*>(lo));
  NSAC_CHECK_LAUNCH("nsac_add_relu_nhwc");
  return NSAC_OK;
}
