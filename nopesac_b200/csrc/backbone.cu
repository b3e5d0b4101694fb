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

// ---- 8-bit image path of the stem (what the reference's data loader delivers: uint8 RGB, siamese_planeTR.py:534-542) ---------
// Raw pixel values 0..255 are exact in fp16, so the im2col matrix needs ONE plane (no lo plane) and the GEMM two passes
// (hi.hi + hi.lo); the normalisation (x - mean) / std is folded into the weights (w / std) and the bias (- sum w mean / std).
// Out-of-image taps are written as 0, which is only right in normalised space: the <= 2 output rows / columns per border whose
// window leaves the image are recomputed exactly by stem_border_fix_kernel afterwards.
// One CTA = one output row (n, yo): its 7 x 3 input rows are staged in shared memory with coalesced loads, then every thread
// emits 16-byte chunks (8 consecutive k of one output pixel).
constexpr int STEM_U8_MAXW = 1024;     // widest image row whose staging buffer (7 rows x 3 channels x (W + 6) fp16) fits 48 KB
constexpr int STEM_U8_THREADS = 256, STEM_U8_LANES = 10;   // 24 k-chunks (8 k each) x 10 pixel lanes (16 threads only help staging)
__global__ void __launch_bounds__(STEM_U8_THREADS)
stem_im2col_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int Ho, int Wo, int border_cols, uint16_t* __restrict__ out) {
  extern __shared__ uint16_t rows[];                   // [7][3][Wp] fp16 bit patterns of the pixel values, x shifted by +3, 0 outside the image
  const int n = blockIdx.x / Ho, yo = blockIdx.x % Ho, Wp = W + 6;
  auto h16 = [](uint32_t v) { return __half_as_ushort(__float2half_rn((float)v)); };     // exact for 0..255
  // coalesced row loads, converted once per input pixel: 4 bytes per thread where rows are 4-byte aligned (W % 4 == 0)
  if ((W & 3) == 0) {
    const int W4 = W >> 2;
    for (int i = threadIdx.x; i < 21 * W4; i += blockDim.x) {
      const int x4 = i % W4, rc = i / W4, ky = rc / 3, c = rc % 3, y = 2 * yo + ky - 3;
      uint32_t v = 0u;
      if (y >= 0 && y < H) v = __ldg(reinterpret_cast<const uint32_t*>(img + (((size_t)n * 3 + c) * H + y) * W) + x4);
      uint16_t* d = rows + rc * Wp + 3 + 4 * x4;
      d[0] = h16(v & 255u); d[1] = h16((v >> 8) & 255u); d[2] = h16((v >> 16) & 255u); d[3] = h16(v >> 24);
    }
    for (int i = threadIdx.x; i < 21 * 6; i += blockDim.x) {          // the 3 + 3 padding columns
      const int rc = i / 6, e = i % 6;
      rows[rc * Wp + (e < 3 ? e : W + e)] = 0;
    }
  } else {
    for (int i = threadIdx.x; i < 21 * Wp; i += blockDim.x) {
      const int xs = i % Wp, rc = i / Wp, ky = rc / 3, c = rc % 3;
      const int y = 2 * yo + ky - 3, x = xs - 3;
      rows[i] = (y >= 0 && y < H && x >= 0 && x < W) ? h16(__ldg(img + (((size_t)n * 3 + c) * H + y) * W + x)) : (uint16_t)0;
    }
  }
  __syncthreads();
  // thread = (k-chunk kc, pixel lane xl): its 8 source offsets are fixed, it walks over the output pixels xl, xl + 10, ...
  const int kc = threadIdx.x % (STEM_KP / 8), xl = threadIdx.x / (STEM_KP / 8);
  int koff[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kc * 8 + e;
    if (k < STEM_K) {
      const int c = k % 3, tap = k / 3, ky = tap / 7, kx = tap - ky * 7;
      koff[e] = (ky * 3 + c) * Wp + kx;
    } else {
      koff[e] = -1;
    }
  }
  const size_t row0 = ((size_t)n * Ho + yo) * Wo;
  for (int xo = xl < STEM_U8_LANES ? xl : Wo; xo < Wo; xo += STEM_U8_LANES) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t v0 = koff[2 * j] >= 0 ? rows[koff[2 * j] + 2 * xo] : 0u, v1 = koff[2 * j + 1] >= 0 ? rows[koff[2 * j + 1] + 2 * xo] : 0u;
      w[j] = v0 | (v1 << 16);
    }
    *reinterpret_cast<uint4*>(out + (row0 + xo) * STEM_KP + kc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (border_cols) {
    // the one-hot class column of the row's border pixels (all pixels of the two first / last rows, four pixels of the others),
    // patched in after the zero padding above has been written by this CTA
    __syncthreads();
    const int rowc = yo == 0 ? 0 : yo == 1 ? 1 : yo == Ho - 2 ? 3 : yo == Ho - 1 ? 4 : 2;
    const int count = rowc != 2 ? Wo : 4;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
      const int xo = rowc != 2 ? i : (i < 2 ? i : Wo - 4 + i);
      const int colc = xo == 0 ? 0 : xo == 1 ? 1 : xo == Wo - 2 ? 3 : xo == Wo - 1 ? 4 : 2, cls = rowc * 5 + colc;
      if (cls != 12) out[(row0 + xo) * STEM_KP + STEM_K + (cls < 12 ? cls : cls - 1)] = 0x3C00;      // 1.0 in fp16
    }
  }
}

// Exact fp32 recomputation of the stem outputs whose 7x7 window leaves the image (zero padding applies to the NORMALISED
// image): out[row, co] = relu(bias[co] + sum_k w[co, k] * (p - mean_c) / std_c over the in-image taps).  w: folded weights
// [64, 147] in (ky, kx, c) order for normalised input, transposed into shared memory ([k][64], 37 KB) once per CTA.  One warp
// per border pixel (grid-stride), lane = 2 output channels.
__global__ void __launch_bounds__(256)
stem_border_fix_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias, int N, int H,
                       int W, int Ho, int Wo, float m0, float m1, float m2, float s0, float s1, float s2, float* __restrict__ out) {
  extern __shared__ float wT[];                        // [147][64]
  for (int i = threadIdx.x; i < 64 * STEM_K; i += blockDim.x) wT[(i % STEM_K) * 64 + i / STEM_K] = __ldg(w + i);
  __syncthreads();
  // border pixels of one image: rows yo < 2 or yo >= Ho - 2 (all xo), plus columns xo < 2 or xo >= Wo - 2 of the other rows
  const int top = Ho < 4 ? Ho : 4, side_rows = Ho - top, per_img = top * Wo + side_rows * 4;
  const int lane = threadIdx.x & 31, warps_per_block = blockDim.x >> 5;
  const float mean[3] = {m0, m1, m2}, istd[3] = {1.f / s0, 1.f / s1, 1.f / s2};
  const float2 b2 = make_float2(__ldg(bias + 2 * lane), __ldg(bias + 2 * lane + 1));
  for (int wi = blockIdx.x * warps_per_block + (threadIdx.x >> 5); wi < N * per_img; wi += gridDim.x * warps_per_block) {
    const int n = wi / per_img, r = wi % per_img;
    int yo, xo;
    if (r < top * Wo) {
      const int t = r / Wo;
      yo = t < 2 ? t : Ho - 4 + t;          // t = 0,1 -> rows 0,1; t = 2,3 -> rows Ho-2, Ho-1
      xo = r % Wo;
    } else {
      const int q = r - top * Wo, t = q & 3;
      yo = 2 + (q >> 2);
      xo = t < 2 ? t : Wo - 4 + t;
    }
    if (yo < 0 || yo >= Ho || xo < 0 || xo >= Wo) continue;
    float a0 = 0.f, a1 = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
      const int y = 2 * yo + ky - 3;
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int x = 2 * xo + kx - 3;
        if (x < 0 || x >= W) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = ((float)__ldg(img + (((size_t)n * 3 + c) * H + y) * W + x) - mean[c]) * istd[c];
          const float2 ww = *reinterpret_cast<const float2*>(wT + ((ky * 7 + kx) * 3 + c) * 64 + 2 * lane);
          a0 = fmaf(ww.x, v, a0);
          a1 = fmaf(ww.y, v, a1);
        }
      }
    }
    float* o = out + (((size_t)n * Ho + yo) * Wo + xo) * 64 + 2 * lane;
    *reinterpret_cast<float2*>(o) = make_float2(fmaxf(a0 + b2.x, 0.f), fmaxf(a1 + b2.y, 0.f));
  }
}

// 3x3 / pad 1 / stride s im2col from NHWC planes to planes [N*Ho*Wo, 9*C] in (ky, kx, c) order: 16-byte copies of both planes
// (one thread = 8 channels of one tap of one output pixel); the strided 3x3 convolutions of res3.0 / res4.0 / res5.0.
__global__ void __launch_bounds__(256)
im2col3x3_from_planes_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int N, int H, int W, int C, int stride,
                             int Ho, int Wo, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  const int C8 = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * 9 * C8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % C8);
    size_t t = idx / C8;
    const int tap = (int)(t % 9);
    const size_t pix = t / 9;
    const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    const int y = yo * stride + tap / 3 - 1, x = xo * stride + tap % 3 - 1;
    uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const size_t src = (((size_t)n * H + y) * W + x) * C + c8 * 8;
      vh = __ldg(reinterpret_cast<const uint4*>(hi + src));
      vl = __ldg(reinterpret_cast<const uint4*>(lo + src));
    }
    const size_t dst = (pix * 9 + tap) * C + c8 * 8;
    *reinterpret_cast<uint4*>(out_hi + dst) = vh;
    *reinterpret_cast<uint4*>(out_lo + dst) = vl;
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

static int stem_im2col_u8_impl(const uint8_t* img, int N, int H, int W, int border_cols, void* out_hi, void* stream) {
  NSAC_REQUIRE(img && out_hi, "nsac_stem_im2col_u8: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 7 && W >= 7 && W <= STEM_U8_MAXW, "nsac_stem_im2col_u8: bad shape N=%d H=%d W=%d", N, H, W);
  if (N == 0) return NSAC_OK;
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  NSAC_REQUIRE((reinterpret_cast<uintptr_t>(img) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0,
               "nsac_stem_im2col_u8: image must be 4-byte aligned, output plane 16-byte aligned");
  const size_t smem = (size_t)21 * (W + 6) * sizeof(uint16_t);
  stem_im2col_u8_kernel<<<N * Ho, STEM_U8_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(img, H, W, Ho, Wo, border_cols,
                                                                                              static_cast<uint16_t*>(out_hi));
  NSAC_CHECK_LAUNCH("nsac_stem_im2col_u8");
  return NSAC_OK;
}

extern "C" int nsac_stem_im2col_u8(const uint8_t* img, int N, int H, int W, void* out_hi, void* stream) {
  return stem_im2col_u8_impl(img, N, H, W, 0, out_hi, stream);
}

// Same, plus 24 one-hot BORDER-CLASS columns (k = 147 .. 170): class = (row class, column class) of the output pixel, each of
// {first, second, interior, second-to-last, last}; interior x interior has no column.  The caller puts, into row 147 + idx of
// the weight matrix, the sum of w * mean / std over the taps that fall outside the image for that class: the GEMM then yields
// the zero-padded convolution of the NORMALISED image for every pixel, and no border pass is needed.
extern "C" int nsac_stem_im2col_u8_cls(const uint8_t* img, int N, int H, int W, void* out_hi, void* stream) {
  NSAC_REQUIRE(H >= 9 && W >= 9, "nsac_stem_im2col_u8_cls: image %dx%d too small for distinct border classes (use nsac_stem_border_fix)", H, W);
  return stem_im2col_u8_impl(img, N, H, W, 1, out_hi, stream);
}

extern "C" int nsac_stem_border_fix(const uint8_t* img, const float* w_folded, const float* bias, int N, int H, int W,
                                    const float* mean3_host, const float* std3_host, float* out, void* stream) {
  NSAC_REQUIRE(img && w_folded && bias && mean3_host && std3_host && out, "nsac_stem_border_fix: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 7 && W >= 7, "nsac_stem_border_fix: bad shape N=%d H=%d W=%d", N, H, W);
  NSAC_REQUIRE(std3_host[0] != 0.f && std3_host[1] != 0.f && std3_host[2] != 0.f, "nsac_stem_border_fix: zero PIXEL_STD");
  if (N == 0) return NSAC_OK;
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const int top = Ho < 4 ? Ho : 4, per_img = top * Wo + (Ho - top) * 4;
  const size_t warps = (size_t)N * per_img;
  const size_t smem_w = (size_t)64 * STEM_K * sizeof(float);
  size_t blocks = (warps + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;          // grid-stride: the 37 KB weight transpose is paid once per CTA
  stem_border_fix_kernel<<<(unsigned)blocks, 256, smem_w, static_cast<cudaStream_t>(stream)>>>(
      img, w_folded, bias, N, H, W, Ho, Wo, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0], std3_host[1], std3_host[2], out);
  NSAC_CHECK_LAUNCH("nsac_stem_border_fix");
  return NSAC_OK;
}

extern "C" int nsac_im2col3x3_from_planes(const void* hi, const void* lo, int N, int H, int W, int C, int stride, void* out_hi,
                                          void* out_lo, void* stream) {
  NSAC_REQUIRE(hi && lo && out_hi && out_lo, "nsac_im2col3x3_from_planes: null pointer");
  NSAC_REQUIRE(N >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && (stride == 1 || stride == 2),
               "nsac_im2col3x3_from_planes: bad shape (C %% 8 == 0, stride 1 or 2)");
  if (N == 0) return NSAC_OK;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  im2col3x3_from_planes_kernel<<<grid_for((size_t)N * Ho * Wo * 9 * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(hi), static_cast<const uint16_t*>(lo), N, H, W, C, stride, Ho, Wo, static_cast<uint16_t*>(out_hi),
      static_cast<uint16_t*>(out_lo));
  NSAC_CHECK_LAUNCH("nsac_im2col3x3_from_planes");
  return NSAC_OK;
}
